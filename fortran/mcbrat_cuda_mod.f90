! mcbrat_cuda_mod.f90 -- ISO_C_BINDING interface to libmcbrat_cuda.so (include/mcbrat_cuda.h).
!
! This is the reference-side binding a maintainer adds to MCBRaT3D: the module below is `use`d
! by Integrators/monteCarloRadiativeTransfer.f95, whose public procedures keep their signatures
! (INT:129-132, 209-218, 845-865, 1046-1073, 1486) while their bodies call these routines.
! NOTE: no Fortran compiler exists in the build image of this repository, so this file is
! reviewed but not compiled here; tests exercise the identical C ABI through ctypes.
module mcbrat_cuda
  use, intrinsic :: iso_c_binding
  implicit none
  private

  integer(c_int), parameter, public :: MCB_ARITH_FAST = 0, MCB_ARITH_REFERENCE = 1

  type, bind(C), public :: mcb_options
    integer(c_int32_t) :: useRayTracing
    integer(c_int32_t) :: useRussianRoulette
    real(c_float)      :: russianRouletteW
    integer(c_int32_t) :: useRussianRouletteForIntensity
    real(c_float)      :: zetaMin
    integer(c_int32_t) :: useHybridPhaseFunsForIntenCalcs
    integer(c_int32_t) :: numOrdersOrigPhaseFunIntenCalcs
    integer(c_int32_t) :: limitIntensityContributions
    real(c_float)      :: maxIntensityContribution
    real(c_float)      :: LW_flag
    integer(c_int32_t) :: arithmetic
    ! measurement knobs (0 = the library's own choice), see include/mcbrat_cuda.h
    integer(c_int32_t) :: tuneKernel, tuneLayout, tuneBlocksPerSM, tuneParkThreshold, tuneLeCarry, tuneExtMask, tuneBurst
    integer(c_int32_t) :: tuneLeap, tuneLeapLanes
    integer(c_int32_t) :: reserved(1)
  end type mcb_options
  integer(c_int32_t), parameter, public :: MCB_KERNEL_PARK = 1, MCB_KERNEL_POOL = 2
  integer(c_int32_t), parameter, public :: MCB_LAYOUT_LINEAR = 1, MCB_LAYOUT_BRICKS = 2

  ! event counters of the last batch (mcb_get_counters)
  type, bind(C), public :: mcb_counters
    integer(c_int64_t) :: photons, crossings, scatters, surfaceHits, topExits, bad, leRays, leCrossings, rouletteKills
    integer(c_int64_t) :: surfaceKills, leaps, leapCells
    integer(c_int64_t) :: reserved(4)
  end type mcb_counters

  ! trace record of the fixed-random-number harness (mcb_run_trace)
  type, bind(C), public :: mcb_event
    integer(c_int32_t) :: photon, kind, ix, iy, iz, component, phaseIndex, angleIndex, order, nrn
    real(c_float)      :: weight, tau
    real(c_double)     :: path, x, y, z
    real(c_float)      :: dir(3)
    integer(c_int32_t) :: pad
  end type mcb_event

  ! one wavelength's description of an optical component for mcb_assemble_optics (read_SSPTable OPT:200-246)
  integer(c_int32_t), parameter, public :: MCB_COMP_VOLEXT = 0, MCB_COMP_ABSXSEC = 1, MCB_COMP_PROFILE = 2
  type, bind(C), public :: mcb_component
    integer(c_int32_t) :: kind, physIndex, nTable, zLevelBase
    type(c_ptr)        :: key        ! real(c_float)(nReff)        -- c_loc(key)
    type(c_ptr)        :: ext        ! real(c_double)(nTable)
    type(c_ptr)        :: ssa        ! real(c_double)(nTable)
    type(c_ptr)        :: phaseIdx   ! integer(c_int32_t)(nTable)
  end type mcb_component

  public :: mcb_create, mcb_destroy, mcb_last_error, mcb_set_grid, mcb_set_optics, mcb_set_inverse_table, &
            mcb_set_forward_table, mcb_set_views, mcb_default_options, mcb_set_options,                  &
            mcb_set_solar_source, mcb_set_thermal_source, mcb_run_batch, mcb_get_results,                &
            mcb_tally_buffer, mcb_synchronize, mcb_status_to_message,                                    &
            mcb_set_physical, mcb_assemble_optics, mcb_get_optics, mcb_build_inverse_table,              &
            mcb_build_thermal_source, mcb_frequency_distribution, mcb_accumulate_batch,                  &
            mcb_stats_reset, mcb_run_batches, mcb_stats_buffer, mcb_get_statistics,                       &
            mcb_version, mcb_set_stream, mcb_build_forward_table, mcb_get_inverse_table, mcb_get_forward_table,  &
            mcb_build_inverse_table_legendre, mcb_build_forward_table_general,                           &
            mcb_get_thermal_source, mcb_last_batch_ms, mcb_get_counters, mcb_get_raw_tallies, mcb_run_trace, &
            mcb_debug_philox, mcb_debug_gather_probe, mcb_debug_distance_map,                            &
            mcb_comm_unique_id, mcb_comm_init, mcb_comm_info, mcb_reduce_tallies, mcb_reduce_statistics, mcb_comm_destroy

  interface
    integer(c_int) function mcb_create(device, handle) bind(C, name="mcb_create")
      import :: c_int, c_ptr
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: handle
    end function
    integer(c_int) function mcb_destroy(handle) bind(C, name="mcb_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function mcb_last_error(handle, buf, len) bind(C, name="mcb_last_error")
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: handle
      character(kind=c_char), intent(out) :: buf(*)
      integer(c_int), value :: len
    end function
    integer(c_int) function mcb_synchronize(handle) bind(C, name="mcb_synchronize")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function mcb_set_grid(handle, nx, ny, nz, xEdges, yEdges, zEdges) bind(C, name="mcb_set_grid")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: nx, ny, nz
      real(c_double), intent(in) :: xEdges(*), yEdges(*), zEdges(*)
    end function
    integer(c_int) function mcb_set_optics(handle, nc, totalExt, cumExt, ssa, phaseIdx, albedo) &
        bind(C, name="mcb_set_optics")
      import :: c_int, c_ptr, c_double, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: nc
      real(c_double), intent(in) :: totalExt(*), cumExt(*), ssa(*)
      integer(c_int32_t), intent(in) :: phaseIdx(*)
      real(c_double), value :: albedo
    end function
    integer(c_int) function mcb_set_inverse_table(handle, comp, nS, nE, T) bind(C, name="mcb_set_inverse_table")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: handle
      integer(c_int), value :: comp, nS, nE
      real(c_float), intent(in) :: T(*)
    end function
    integer(c_int) function mcb_set_forward_table(handle, comp, nS, nE, P, Porig) bind(C, name="mcb_set_forward_table")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: handle
      integer(c_int), value :: comp, nS, nE
      real(c_float), intent(in) :: P(*), Porig(*)
    end function
    integer(c_int) function mcb_set_views(handle, nDir, dirCos) bind(C, name="mcb_set_views")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: handle
      integer(c_int), value :: nDir
      real(c_float), intent(in) :: dirCos(*)
    end function
    subroutine mcb_default_options(o) bind(C, name="mcb_default_options")
      import :: mcb_options
      type(mcb_options), intent(out) :: o
    end subroutine
    integer(c_int) function mcb_set_options(handle, o) bind(C, name="mcb_set_options")
      import :: c_int, c_ptr, mcb_options
      type(c_ptr), value :: handle
      type(mcb_options), intent(in) :: o
    end function
    integer(c_int) function mcb_set_solar_source(handle, solarMu, solarAzimuthDeg) bind(C, name="mcb_set_solar_source")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: handle
      real(c_float), value :: solarMu, solarAzimuthDeg
    end function
    integer(c_int) function mcb_set_thermal_source(handle, fracAtmsPower, voxelCDF) bind(C, name="mcb_set_thermal_source")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), value :: fracAtmsPower
      real(c_double), intent(in) :: voxelCDF(*)
    end function
    integer(c_int) function mcb_run_batch(handle, nPhotons, seed, firstPhotonId, nProcessed) bind(C, name="mcb_run_batch")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: nPhotons, seed, firstPhotonId
      integer(c_int64_t), intent(out) :: nProcessed
    end function
    integer(c_int) function mcb_get_results(handle, nPhotonsNormalise, fluxUp, fluxDown, fluxAbsorbed, &
                                            volumeAbsorption, intensity, intensityByComponent) bind(C, name="mcb_get_results")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: nPhotonsNormalise
      type(c_ptr), value :: fluxUp, fluxDown, fluxAbsorbed, volumeAbsorption, intensity, intensityByComponent
    end function
    integer(c_int) function mcb_tally_buffer(handle, devicePtr, nDoubles) bind(C, name="mcb_tally_buffer")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: handle
      type(c_ptr), intent(out) :: devicePtr
      integer(c_int64_t), intent(out) :: nDoubles
    end function
    ! ---- read_SSPTable on the device: commonDomain once (OPT:63-75), then one call per wavelength ----
    integer(c_int) function mcb_set_physical(handle, nPhys, massConc, Reff, numConc) bind(C, name="mcb_set_physical")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: nPhys
      real(c_double), intent(in) :: massConc(*), Reff(*)      ! commonD%massConc(comp,x,y,z), commonD%Reff
      real(c_double), intent(in) :: numConc(*)                ! commonD%numConc(1,1,:)
    end function
    integer(c_int) function mcb_assemble_optics(handle, nc, comps, setup, albedo) bind(C, name="mcb_assemble_optics")
      import :: c_int, c_ptr, c_double, mcb_component
      type(c_ptr), value :: handle
      integer(c_int), value :: nc, setup
      type(mcb_component), intent(in) :: comps(*)
      real(c_double), value :: albedo
    end function
    integer(c_int) function mcb_get_optics(handle, totalExt, cumExt, ssa, phaseIdx) bind(C, name="mcb_get_optics")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle, totalExt, cumExt, ssa, phaseIdx     ! c_loc(array) or c_null_ptr
    end function
    integer(c_int) function mcb_build_inverse_table(handle, comp, nS, nE, nAngles, mus, values) &
        bind(C, name="mcb_build_inverse_table")
      import :: c_int, c_ptr, c_float, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: comp, nS, nE
      integer(c_int32_t), intent(in) :: nAngles(*)
      real(c_float), intent(in) :: mus(*), values(*)
    end function
    ! ---- emission_weighting (EMI:424-550) and getFrequencyDistr (EMI:552-573) on the device ----
    integer(c_int) function mcb_build_thermal_source(handle, temps, lambda_um, surfaceTemp, fracAtmsPower, totalFlux) &
        bind(C, name="mcb_build_thermal_source")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      type(c_ptr), value :: temps            ! c_loc(temps(1,1,1)), or c_null_ptr: the temperatures staged before
      real(c_double), value :: lambda_um, surfaceTemp
      real(c_double), intent(out) :: fracAtmsPower, totalFlux
    end function
    integer(c_int) function mcb_frequency_distribution(handle, nLambda, cdf, totalPhotons, seed, distribution) &
        bind(C, name="mcb_frequency_distribution")
      import :: c_int, c_ptr, c_double, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int), value :: nLambda
      real(c_double), intent(in) :: cdf(*)
      integer(c_int64_t), value :: totalPhotons, seed
      integer(c_int64_t), intent(out) :: distribution(*)
    end function
    ! ---- the driver's batch loop and statistics (DRV:949-1052, 1188-1228) on the device ----
    integer(c_int) function mcb_accumulate_batch(handle, nPhotons, seed, firstPhotonId, nProcessed) &
        bind(C, name="mcb_accumulate_batch")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: nPhotons, seed, firstPhotonId
      integer(c_int64_t), intent(out) :: nProcessed
    end function
    integer(c_int) function mcb_stats_reset(handle) bind(C, name="mcb_stats_reset")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function mcb_run_batches(handle, numBatches, photonsPerBatch, seed, firstPhotonId, nProcessed) &
        bind(C, name="mcb_run_batches")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: numBatches, photonsPerBatch, seed, firstPhotonId
      integer(c_int64_t), intent(out) :: nProcessed
    end function
    integer(c_int) function mcb_stats_buffer(handle, devicePtr, nDoubles) bind(C, name="mcb_stats_buffer")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: handle
      type(c_ptr), intent(out) :: devicePtr
      integer(c_int64_t), intent(out) :: nDoubles
    end function
    integer(c_int) function mcb_get_statistics(handle, solarFlux, meanFluxStats, fluxUpStats, fluxDownStats, &
        fluxAbsorbedStats, absorbedProfileStats, absorbedVolumeStats, radianceStats, totalNumPhotons, batchesCompleted) &
        bind(C, name="mcb_get_statistics")
      import :: c_int, c_ptr, c_double, c_int64_t
      type(c_ptr), value :: handle
      real(c_double), value :: solarFlux
      type(c_ptr), value :: meanFluxStats, fluxUpStats, fluxDownStats, fluxAbsorbedStats, absorbedProfileStats, &
                            absorbedVolumeStats, radianceStats        ! c_loc(stats array) or c_null_ptr
      integer(c_int64_t), intent(out) :: totalNumPhotons, batchesCompleted
    end function
    ! ---- the rest of the header: version, stream, table builders / read-back, timing, counters, trace, aids ----
    integer(c_int) function mcb_version() bind(C, name="mcb_version")
      import :: c_int
    end function
    integer(c_int) function mcb_set_stream(handle, cudaStream) bind(C, name="mcb_set_stream")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle, cudaStream                 ! a cudaStream_t, or c_null_ptr for the handle's own
    end function
    integer(c_int) function mcb_build_forward_table(handle, comp, nS, nE, nCoef, coefs) bind(C, name="mcb_build_forward_table")
      import :: c_int, c_ptr, c_float, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: comp, nS, nE
      integer(c_int32_t), intent(in) :: nCoef(*)
      real(c_float), intent(in) :: coefs(*)
    end function
    ! computeInversePhaseFuncTable (INV:66-174) for Legendre-stored tables, Lobatto nodes and all, in HBM
    integer(c_int) function mcb_build_inverse_table_legendre(handle, comp, nS, nE, nCoef, coefs) &
        bind(C, name="mcb_build_inverse_table_legendre")
      import :: c_int, c_ptr, c_float, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: comp, nS, nE
      integer(c_int32_t), intent(in) :: nCoef(*)
      real(c_float), intent(in) :: coefs(*)
    end function
    ! tabulateForwardPhaseFunctions (OPT:1872-1934) for either storage kind, with the hybrid peak (OPT:1936-2050) if asked
    integer(c_int) function mcb_build_forward_table_general(handle, comp, nS, nE, nCoef, coefs, nAngles, angles, values, &
        hybridWidthDeg) bind(C, name="mcb_build_forward_table_general")
      import :: c_int, c_ptr, c_float, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: comp, nS, nE
      integer(c_int32_t), intent(in) :: nCoef(*), nAngles(*)
      real(c_float), intent(in) :: coefs(*), angles(*), values(*)
      real(c_float), value :: hybridWidthDeg
    end function
    integer(c_int) function mcb_get_inverse_table(handle, comp, T, nFloats) bind(C, name="mcb_get_inverse_table")
      import :: c_int, c_ptr, c_float, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int), value :: comp
      real(c_float), intent(out) :: T(*)
      integer(c_int64_t), value :: nFloats
    end function
    integer(c_int) function mcb_get_forward_table(handle, comp, T, nFloats) bind(C, name="mcb_get_forward_table")
      import :: c_int, c_ptr, c_float, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int), value :: comp
      real(c_float), intent(out) :: T(*)
      integer(c_int64_t), value :: nFloats
    end function
    integer(c_int) function mcb_get_thermal_source(handle, fracAtmsPower, voxelCDF, nDoubles) bind(C, name="mcb_get_thermal_source")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: handle
      type(c_ptr), value :: fracAtmsPower, voxelCDF            ! c_loc(real(c_double)) or c_null_ptr
      integer(c_int64_t), value :: nDoubles
    end function
    integer(c_int) function mcb_last_batch_ms(handle, ms) bind(C, name="mcb_last_batch_ms")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: handle
      real(c_float), intent(out) :: ms
    end function
    integer(c_int) function mcb_get_counters(handle, c) bind(C, name="mcb_get_counters")
      import :: c_int, c_ptr, mcb_counters
      type(c_ptr), value :: handle
      type(mcb_counters), intent(out) :: c
    end function
    integer(c_int) function mcb_get_raw_tallies(handle, out, nDoubles) bind(C, name="mcb_get_raw_tallies")
      import :: c_int, c_ptr, c_double, c_int64_t
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: out(*)
      integer(c_int64_t), value :: nDoubles
    end function
    integer(c_int) function mcb_run_trace(handle, nPhotons, rn, rnStride, maxEventsPerPhoton, events, eventCap, nEvents) &
        bind(C, name="mcb_run_trace")
      import :: c_int, c_ptr, c_float, c_int32_t, c_int64_t, mcb_event
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: nPhotons, rnStride, eventCap
      real(c_float), intent(in) :: rn(*)
      integer(c_int32_t), value :: maxEventsPerPhoton
      type(mcb_event), intent(out) :: events(*)
      integer(c_int64_t), intent(out) :: nEvents
    end function
    integer(c_int) function mcb_debug_philox(handle, seed, photon, n, out) bind(C, name="mcb_debug_philox")
      import :: c_int, c_ptr, c_int32_t, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: seed, photon
      integer(c_int), value :: n
      integer(c_int32_t), intent(out) :: out(*)
    end function
    integer(c_int) function mcb_debug_gather_probe(handle, bytes, loadsInFlight, blocksPerSM, iterations, gathersPerSecond) &
        bind(C, name="mcb_debug_gather_probe")
      import :: c_int, c_ptr, c_int64_t, c_double
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: bytes
      integer(c_int), value :: loadsInFlight, blocksPerSM, iterations
      real(c_double), intent(out) :: gathersPerSecond
    end function
    integer(c_int) function mcb_debug_distance_map(handle, out, nBytes) bind(C, name="mcb_debug_distance_map")
      import :: c_int, c_ptr, c_int64_t, c_int8_t
      type(c_ptr), value :: handle
      integer(c_int8_t), intent(out) :: out(*)
      integer(c_int64_t), value :: nBytes
    end function
    ! ---- multipleProcesses (MPIW:29-251) over NCCL: one process per GPU, one reduce at the end (DRV:1151-1166) ----
    integer(c_int) function mcb_comm_unique_id(id128) bind(C, name="mcb_comm_unique_id")
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id128(128)        ! rank 0 creates it; MPI_BCAST carries it to the others
    end function
    integer(c_int) function mcb_comm_init(handle, nranks, rank, id128) bind(C, name="mcb_comm_init")
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: handle
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id128(128)
    end function
    integer(c_int) function mcb_comm_info(handle, nranks, rank, ncclVersion) bind(C, name="mcb_comm_info")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), intent(out) :: nranks, rank, ncclVersion
    end function
    integer(c_int) function mcb_reduce_tallies(handle, root) bind(C, name="mcb_reduce_tallies")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: root                            ! >= 0: MPI_REDUCE to that rank; < 0: all-reduce
    end function
    integer(c_int) function mcb_reduce_statistics(handle, root) bind(C, name="mcb_reduce_statistics")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: root
    end function
    integer(c_int) function mcb_comm_destroy(handle) bind(C, name="mcb_comm_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
  end interface

contains

  ! Non-zero return code -> the text handed to setStateToFailure(status, ...) (ErrorMessages.f95:225)
  function mcb_status_to_message(handle) result(msg)
    type(c_ptr), intent(in) :: handle
    character(len=512) :: msg
    character(kind=c_char) :: buf(512)
    integer :: i, rc
    msg = ""
    rc = mcb_last_error(handle, buf, 512_c_int)
    do i = 1, 512
      if (buf(i) == c_null_char) exit
      msg(i:i) = buf(i)
    end do
  end function mcb_status_to_message

end module mcbrat_cuda

! ---------------------------------------------------------------------------------------------
! How the reference's procedures call it (sketch of the replaced bodies; signatures unchanged):
!
!   type integrator            ! INT:40-117 gains one member
!     type(c_ptr) :: gpu = c_null_ptr
!     integer(8)  :: nextPhotonId = 0
!   end type
!
!   function new_Integrator(atmosphere, status) result(new)                      ! INT:129-201
!     ... existing getInfo_Domain calls for xPosition/yPosition/zPosition ...
!     if (mcb_create(0_c_int, new%gpu) /= 0) call setStateToFailure(status, "new_Integrator: no CUDA device")
!     if (mcb_set_grid(new%gpu, numX, numY, numZ, new%xPosition, new%yPosition, new%zPosition) /= 0) &
!       call setStateToFailure(status, "new_Integrator: " // trim(mcb_status_to_message(new%gpu)))
!   end function
!
!   subroutine computeRadiativeTransfer(thisIntegrator, thisDomain, randomNumbers, incomingPhotons, &
!                                       numPhotonsPerBatch, numPhotonsProcessed, status)              ! INT:209-218
!     call tabulateInversePhaseFunctions(thisDomain, thisIntegrator%minInverseTableSize, status)      ! INT:280 (host)
!     call getInfo_Domain(thisDomain, albedo=albedo, totalExt=totalExt, cumExt=cumExt, ssa=ssa, &
!                         phaseFuncI=phaseFuncI, inversePhaseFuncs=inversePhaseFuncs, status=status)  ! INT:441-443
!     rc = mcb_set_optics(thisIntegrator%gpu, numComps, totalExt, cumExt, ssa, phaseFuncI, albedo)    ! once per domain
!     do i = 1, numComps
!       rc = mcb_set_inverse_table(thisIntegrator%gpu, i, size(inversePhaseFuncs(i)%values, 1), &
!                                  size(inversePhaseFuncs(i)%values, 2), inversePhaseFuncs(i)%values)
!     end do
!     rc = mcb_set_solar_source(thisIntegrator%gpu, solarMu, solarAzimuth)     ! or mcb_set_thermal_source
!     rc = mcb_run_batch(thisIntegrator%gpu, min(numPhotonsPerBatch, photonsLeft), seed, &
!                        thisIntegrator%nextPhotonId, numPhotonsProcessed)
!     if (rc /= 0) call setStateToFailure(status, "computeRadiativeTransfer: " // &
!                                         trim(mcb_status_to_message(thisIntegrator%gpu)))
!     thisIntegrator%nextPhotonId = thisIntegrator%nextPhotonId + numPhotonsProcessed
!     rc = mcb_get_results(thisIntegrator%gpu, 0_c_int64_t, c_loc(thisIntegrator%fluxUp), c_loc(thisIntegrator%fluxDown), &
!                          c_loc(thisIntegrator%fluxAbsorbed), c_loc(thisIntegrator%volumeAbsorption), &
!                          c_loc(thisIntegrator%intensity), c_loc(thisIntegrator%intensityByComponent))
!     ! reportResults (INT:845-1042) is unchanged: it copies out of thisIntegrator%fluxUp etc.
!   end subroutine
!
!   ! ---- multi-GPU (one MPI rank per GPU; the domain replicated; photons split by global photon id) ----
!   ! src/multipleProcesses_mpi.f95 keeps MPI for process start-up; the tallies no longer travel through
!   ! sumAcrossProcesses (MPIW:70-251) but through ONE NCCL reduce on the device:
!   !
!   !   call initializeProcesses(numProcs, thisProc)                              ! MPIW:29-52, unchanged (MPI_INIT ...)
!   !   rc = mcb_create(int(mod(thisProc, gpusPerNode), c_int), mcIntegrator%gpu) ! in new_Integrator
!   !   if (MasterProc) rc = mcb_comm_unique_id(ncclId)
!   !   call MPI_BCAST(ncclId, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
!   !   rc = mcb_comm_init(mcIntegrator%gpu, int(numProcs, c_int), int(thisProc, c_int), ncclId)
!   !   ... every rank, the master included, runs its block of batches:
!   !   rc = mcb_run_batches(mcIntegrator%gpu, batchesPerProc, numPhotonsPerBatch, seed, &
!   !                        thisProc * batchesPerProc * numPhotonsPerBatch, nDone)
!   !   ... DRV:1151-1166, nine sumAcrossProcesses calls on host arrays, become
!   !   rc = mcb_reduce_statistics(mcIntegrator%gpu, 0_c_int)
!   !   if (MasterProc) rc = mcb_get_statistics(mcIntegrator%gpu, solarFlux, c_loc(meanFluxStats), c_loc(fluxUpStats), ...)
!   !   rc = mcb_comm_destroy(mcIntegrator%gpu); call finalizeProcesses()         ! MPIW:62-68
