#!/bin/bash
# Is there ANY Fortran front end (or MPI / netCDF-Fortran) on the GPU box?  oracle/_ref needs one to compile the
# unmodified reference sources.  The log is committed under profiles/ (VERDICT r1, next #2).
out=${1:-gpurun_out/r02_compiler_probe.log}
{
  echo "== host: $(uname -a)"
  echo "== nproc: $(nproc)"
  for c in gfortran gfortran-9 gfortran-10 gfortran-11 gfortran-12 gfortran-13 gfortran-14 flang flang-new flang-18 lfortran nvfortran pgfortran pgf90 ifort ifx f77 f95 g95 f2c xlf mpif90 mpifort mpicc mpirun mpiexec nc-config nf-config ncdump h5fc; do
    p=$(command -v $c 2>/dev/null); echo "$c: ${p:-absent}"
  done
  echo "== find (names containing fortran|flang|f951|nvfortran, outside /proc):"
  find / -xdev \( -name 'f951' -o -name '*gfortran*' -o -name 'flang*' -o -name 'nvfortran*' -o -name 'lfortran*' -o -name 'netcdf.mod' -o -name 'mpif.h' \) 2>/dev/null | grep -v '^/proc' | head -40
  echo "== /opt/nvidia/hpc_sdk: $(ls /opt/nvidia/hpc_sdk 2>&1 | head -3)"
  echo "== gcc: $(gcc --version | head -1); languages: $(gcc -v 2>&1 | grep -o 'enable-languages=[^ ]*')"
  echo "== python numpy.f2py compilers:"; python -c "import numpy.f2py, shutil; print('f2py present; gfortran on PATH:', shutil.which('gfortran'))" 2>&1
} > "$out" 2>&1
cat "$out"
