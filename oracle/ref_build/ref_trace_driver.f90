! ref_trace_driver.f90 -- drives the UNMODIFIED reference modules of MCBRaT3D one photon at a time.
!
! TEST INFRASTRUCTURE (oracle/).  Compiled by oracle/ref_build/Makefile together with the reference's own sources
! (Integrators/monteCarloRadiativeTransfer.f95, src/opticalProperties.f95, src/monteCarloIllumination.f95,
! src/emissionAndBroadBandWeights.f95, src/RandomNumbersForMC.f95, src/numericUtilities.f95,
! src/inversePhaseFunctions.f95, src/scatteringPhaseFunctions.f95, src/surfaceProperties.f95, src/ErrorMessages.f95,
! src/characterUtils.f95) where they lie under /root/reference, and the generated netcdf stand-in.
!
! What it does: reads a case file written by make_ref_cases.py (grid, components, phase functions, views, seed), builds
! the domain through the reference's public API exactly as Drivers/monteCarloDriver.f95 does (new_Domain,
! addOpticalComponent, new_Integrator, specifyParameters: DRV:533-597), seeds the reference's own MT19937 with
! (/ iseed, rank, 0 /) (DRV:901) and then
!   mode 0 (fingerprints): runs numBatches batches of photonsPerBatch photons (1 for per-photon fingerprints) through
!           new_PhotonStream + computeRadiativeTransfer + reportResults (DRV:956-1011) and writes every non-zero entry of
!           fluxUp, fluxDown, fluxAbsorbed, volumeAbsorption and intensity after each batch.  With one photon per batch
!           and absorbing media that is the photon's whole history: the cell and the weight of every scattering event,
!           the exit column and weight, and every local-estimate contribution by pixel.  The C oracle, seeded the same
!           way, must reproduce these numbers exactly (tests/test_ref_fixtures.py);
!   mode 1 (timing): runs the batches and prints photons and seconds (bench.py --impl reference, kind "reference").
!
! Case file (stream access, native little endian), in this order:
!   int32  nx, ny, nz, nc, nDir, iseed, rank, mode, nS, useRRforIntensity
!   int64  numBatches, photonsPerBatch
!   real64 albedo;  real32 solarMu, solarAzimuth, zetaMin
!   real64 xEdges(nx+1), yEdges(ny+1), zEdges(nz+1)
!   real32 mus(nDir), phis(nDir)
!   per component: int32 zLevelBase, uniform, nzc, nEntries; per entry: int32 n (> 0: n Legendre coefficients follow as
!     real32; < 0: -n scattering angles then -n values, real32); then extinction, singleScatteringAlbedo (real64) and
!     phaseFunctionIndex (int32), dimensioned (nzc) when uniform /= 0, else (nx, ny, nzc) in Fortran order
! Output file (stream): per batch  int32 batch, int32 photonsProcessed, int32 nNonZero, then nNonZero x (int32 array id
!   1..5, int32 linear index (1-based, Fortran order), real32 value).
program ref_trace_driver
  use ErrorMessages
  use RandomNumbers
  use scatteringPhaseFunctions
  use opticalProperties
  use monteCarloIllumination
  use monteCarloRadiativeTransfer
  implicit none

  character(len=1024) :: caseFile, outFile
  integer(4) :: nx, ny, nz, nc, nDir, iseed, rank, mode, nS, useRR
  integer(8) :: numBatches, photonsPerBatch, numPhotonsProcessed, b, total
  real(8)    :: albedo
  real(4)    :: solarMu, solarAzimuth, zetaMin
  real(4), allocatable :: mus(:), phis(:)
  integer(4) :: zLevelBase, uniform, nzc, nEntries, n, c, e
  real(4), allocatable :: coef(:), ang(:), val(:), key(:)
  real(8), allocatable :: ext1(:), ssa1(:), ext3(:, :, :), ssa3(:, :, :)
  integer(4), allocatable :: idx1(:), idx3(:, :, :)
  type(phaseFunction), allocatable :: pfs(:)
  type(phaseFunctionTable) :: table
  type(commonDomain), target :: commonD
  type(domain) :: thisDomain
  type(integrator) :: mcIntegrator
  type(randomNumberSequence) :: randoms
  type(photonStream) :: incomingPhotons
  type(ErrorMessage) :: status
  real(4), allocatable :: fluxUp(:, :), fluxDown(:, :), fluxAbsorbed(:, :), absorbedVolume(:, :, :), intensity(:, :, :)
  integer :: count0, count1, countRate
  character(len=16) :: compName

  if (command_argument_count() < 2) then
    print *, "usage: ref_trace_driver CASEFILE OUTFILE"
    stop 2
  end if
  call get_command_argument(1, caseFile)
  call get_command_argument(2, outFile)
  open(unit=10, file=trim(caseFile), access="stream", form="unformatted", status="old", action="read")
  read(10) nx, ny, nz, nc, nDir, iseed, rank, mode, nS, useRR
  read(10) numBatches, photonsPerBatch
  read(10) albedo
  read(10) solarMu, solarAzimuth, zetaMin
  allocate(commonD%xPosition(nx + 1), commonD%yPosition(ny + 1), commonD%zPosition(nz + 1), commonD%temps(nx, ny, nz))
  read(10) commonD%xPosition
  read(10) commonD%yPosition
  read(10) commonD%zPosition
  commonD%temps = 0.0_8
  allocate(mus(max(nDir, 1)), phis(max(nDir, 1)))
  if (nDir > 0) then
    read(10) mus(1:nDir)
    read(10) phis(1:nDir)
  end if

  ! ---- the domain, through the reference's own constructors (OPT:455-507, 557-665) ----
  thisDomain = new_Domain(commonD, 0.0_8, 1, 1, albedo, status)
  call check("new_Domain")
  do c = 1, nc
    read(10) zLevelBase, uniform, nzc, nEntries
    allocate(pfs(nEntries), key(nEntries))
    do e = 1, nEntries
      read(10) n
      if (n > 0) then                                    ! Legendre coefficients (SPF:166-227)
        allocate(coef(n))
        read(10) coef
        pfs(e) = new_PhaseFunction(coef, status = status)
        deallocate(coef)
      else                                               ! scattering angle / value pairs (SPF:101-165)
        allocate(ang(-n), val(-n))
        read(10) ang
        read(10) val
        pfs(e) = new_PhaseFunction(ang, val, status = status)
        deallocate(ang, val)
      end if
      call check("new_PhaseFunction")
      key(e) = real(e)
    end do
    table = new_PhaseFunctionTable(pfs, key = key, status = status)
    call check("new_PhaseFunctionTable")
    write(compName, "(A,I0)") "component", c
    if (uniform /= 0) then
      allocate(ext1(nzc), ssa1(nzc), idx1(nzc))
      read(10) ext1
      read(10) ssa1
      read(10) idx1
      call addOpticalComponent(thisDomain, trim(compName), ext1, ssa1, idx1, table, zLevelBase = zLevelBase, status = status)
      deallocate(ext1, ssa1, idx1)
    else
      allocate(ext3(nx, ny, nzc), ssa3(nx, ny, nzc), idx3(nx, ny, nzc))
      read(10) ext3
      read(10) ssa3
      read(10) idx3
      call addOpticalComponent(thisDomain, trim(compName), ext3, ssa3, idx3, table, zLevelBase = zLevelBase, status = status)
      deallocate(ext3, ssa3, idx3)
    end if
    call check("addOpticalComponent")
    do e = 1, nEntries
      call finalize_PhaseFunction(pfs(e))
    end do
    deallocate(pfs, key)
  end do
  close(10)

  ! ---- the integrator, as the driver sets it up (DRV:533-597) ----
  mcIntegrator = new_Integrator(thisDomain, status = status)
  call check("new_Integrator")
  call specifyParameters(mcIntegrator, minInverseTableSize = nS, LW_flag = -1.0, status = status)
  call check("specifyParameters")
  if (nDir > 0) then
    call specifyParameters(mcIntegrator, minForwardTableSize = nS, intensityMus = mus(1:nDir), intensityPhis = phis(1:nDir), &
                           computeIntensity = .true., numComps = nc, status = status)
    call check("specifyParameters (intensity)")
  end if
  call specifyParameters(mcIntegrator, useRayTracing = .true., useRussianRoulette = .true., status = status)
  call check("specifyParameters (algorithm)")
  if (nDir > 0) then
    call specifyParameters(mcIntegrator, useHybridPhaseFunsForIntenCalcs = .false., numOrdersOrigPhaseFunIntenCalcs = 0, &
                           useRussianRouletteForIntensity = (useRR /= 0), zetaMin = zetaMin,                          &
                           limitIntensityContributions = .false., numComps = nc, status = status)
    call check("specifyParameters (intensity algorithm)")
  end if
  randoms = new_RandomNumberSequence(seed = (/ iseed, rank, 0 /))                        ! DRV:901

  allocate(fluxUp(nx, ny), fluxDown(nx, ny), fluxAbsorbed(nx, ny), absorbedVolume(nx, ny, nz), intensity(nx, ny, max(nDir, 1)))
  if (mode == 0) open(unit=20, file=trim(outFile), access="stream", form="unformatted", status="replace", action="write")
  total = 0
  call system_clock(count0, countRate)
  do b = 1, numBatches                                                                    ! DRV:956-1011
    incomingPhotons = new_PhotonStream(solarMu, solarAzimuth, numberOfPhotons = photonsPerBatch, &
                                       randomNumbers = randoms, status = status)
    call check("new_PhotonStream")
    call computeRadiativeTransfer(mcIntegrator, thisDomain, randoms, incomingPhotons, photonsPerBatch, &
                                  numPhotonsProcessed, status)
    call check("computeRadiativeTransfer")
    call finalize_PhotonStream(incomingPhotons)
    total = total + numPhotonsProcessed
    if (mode == 0) then
      if (nDir > 0) then
        call reportResults(mcIntegrator, fluxUp = fluxUp, fluxDown = fluxDown, fluxAbsorbed = fluxAbsorbed, &
                           volumeAbsorption = absorbedVolume, intensity = intensity(:, :, 1:nDir), status = status)
      else
        call reportResults(mcIntegrator, fluxUp = fluxUp, fluxDown = fluxDown, fluxAbsorbed = fluxAbsorbed, &
                           volumeAbsorption = absorbedVolume, status = status)
        intensity = 0.0
      end if
      call check("reportResults")
      call dump(int(b, 4), int(numPhotonsProcessed, 4))
    end if
  end do
  call system_clock(count1)
  if (mode == 0) then
    close(20)
  else
    open(unit=20, file=trim(outFile), status="replace", action="write")
    write(20, "(I0,1X,ES16.8)") total, real(count1 - count0, 8) / real(countRate, 8)
    close(20)
  end if
  print "(A,I0,A,ES12.4,A)", "ref_trace_driver: ", total, " photons in ", real(count1 - count0, 8) / real(countRate, 8), " s"

contains

  subroutine check(where)
    character(len=*), intent(in) :: where
    if (stateIsFailure(status)) then
      print *, "ref_trace_driver: failure in ", where, ": ", trim(getCurrentMessage(status))
      stop 1
    end if
  end subroutine check

  ! every non-zero entry of the five result arrays: (array id, 1-based linear index in Fortran order, value)
  subroutine dump(batch, processed)
    integer(4), intent(in) :: batch, processed
    integer(4) :: nnz, i, j, k
    nnz = count(fluxUp /= 0.0) + count(fluxDown /= 0.0) + count(fluxAbsorbed /= 0.0) + count(absorbedVolume /= 0.0)
    if (nDir > 0) nnz = nnz + count(intensity(:, :, 1:nDir) /= 0.0)
    write(20) batch, processed, nnz
    do j = 1, ny
      do i = 1, nx
        if (fluxUp(i, j) /= 0.0) write(20) 1_4, int(i + nx * (j - 1), 4), fluxUp(i, j)
      end do
    end do
    do j = 1, ny
      do i = 1, nx
        if (fluxDown(i, j) /= 0.0) write(20) 2_4, int(i + nx * (j - 1), 4), fluxDown(i, j)
      end do
    end do
    do j = 1, ny
      do i = 1, nx
        if (fluxAbsorbed(i, j) /= 0.0) write(20) 3_4, int(i + nx * (j - 1), 4), fluxAbsorbed(i, j)
      end do
    end do
    do k = 1, nz
      do j = 1, ny
        do i = 1, nx
          if (absorbedVolume(i, j, k) /= 0.0) write(20) 4_4, int(i + nx * ((j - 1) + ny * (k - 1)), 4), absorbedVolume(i, j, k)
        end do
      end do
    end do
    if (nDir > 0) then
      do k = 1, nDir
        do j = 1, ny
          do i = 1, nx
            if (intensity(i, j, k) /= 0.0) write(20) 5_4, int(i + nx * ((j - 1) + ny * (k - 1)), 4), intensity(i, j, k)
          end do
        end do
      end do
    end if
  end subroutine dump

end program ref_trace_driver
