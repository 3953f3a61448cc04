! ISO_C_BINDING interface to libeqdyna_b200.so (include/eqdyna_b200.h): the B200 step
! library that replaces the body of `subroutine driver` (src/driver.f90:3-36).
!
! Not compiled in the development image (no gfortran there); kept in lock-step with
! the C header by tests/test_host_and_abi.py::test_fortran_interface_covers_the_abi.
!
! Host-side changes needed in EQdyna (see INTEGRATION.md):
!   * the globalvar arrays handed over below need the TARGET attribute
!     (src/globalvar.f90:81-101), a one-word change per declaration;
!   * src/makefile: add eqdyna_cuda_iface.o driver_cuda.o, link -leqdyna_b200 -lcudart.
module eqdyna_cuda_iface
    use, intrinsic :: iso_c_binding
    implicit none

    integer(c_int), parameter :: EQD_OK = 0, EQD_ERR_NAN = 1, EQD_ERR_DAMP = 2, EQD_ERR_CUDA = 3, EQD_ERR_ARG = 4
    integer(c_int32_t), parameter :: EQD_F_DISP = 1, EQD_F_VEL = 2, EQD_F_V1 = 3, EQD_F_FORCE = 4, EQD_F_FRIC = 5, &
        EQD_F_FNFT = 6, EQD_F_PSTRAIN = 7, EQD_F_STRESS = 8, EQD_F_ONFAULT_HIST = 9, EQD_F_OFFFAULT_HIST = 10, &
        EQD_F_HYPO_LOG = 11, EQD_F_GM = 12, EQD_F_SRC_EVOL = 13, EQD_F_TPHIST = 14, EQD_F_MASS = 15, EQD_F_FNMS = 16, &
        EQD_F_ARN = 17

    ! struct eqd_params -- field order is the ABI
    type, bind(C) :: eqd_params
        real(c_double)     :: dt
        integer(c_int32_t) :: nstep, me, npx, npy, npz
        real(c_double)     :: rdampk, rdampm, w, grav, roumax, rhow, gamar, ccosphi, sinphi, tv, kapa_hg, dx
        integer(c_int32_t) :: C_elastic, C_Q, C_hg
        real(c_double)     :: PMLb(8)
        integer(c_int32_t) :: nPML
        real(c_double)     :: R, vmaxPML
        integer(c_int32_t) :: friclaw, C_nuclea, nucfault, TPV, insertFaultType, ntotft
        real(c_double)     :: nucR, nucT, nucRuptVel, nucdtau0, xsource, ysource, zsource, slipRateThres, tol, fric_tp_h
        integer(c_int32_t) :: outputGroundMotion, reserved_i(7)
        real(c_double)     :: reserved_d(8)
    end type eqd_params

    interface
        integer(c_int) function eqd_create(p, device, handle) bind(C, name='eqd_create')
            import :: c_int, c_ptr, eqd_params
            type(eqd_params), intent(in) :: p
            integer(c_int), value :: device          ! -1: me modulo the visible device count
            type(c_ptr), intent(out) :: handle
        end function
        integer(c_int) function eqd_destroy(handle) bind(C, name='eqd_destroy')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
        end function
        integer(c_int) function eqd_last_error(handle, buf, n) bind(C, name='eqd_last_error')
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: handle
            character(kind=c_char) :: buf(*)
            integer(c_int), value :: n
        end function
        integer(c_int) function eqd_set_mesh(handle, Nn, Ne, Neq, sizeEq, meshCoor, nodeElemIdRelation, elemTypeArr, &
                numOfDofPerNodeArr, eqNumStartIndexLoc, eqNumIndexArr, stressCompIndexArr, sizeStress) bind(C, name='eqd_set_mesh')
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: Nn, Ne, Neq, sizeEq, sizeStress
            type(c_ptr), value :: meshCoor, nodeElemIdRelation, elemTypeArr, numOfDofPerNodeArr, eqNumStartIndexLoc, &
                                  eqNumIndexArr, stressCompIndexArr
        end function
        integer(c_int) function eqd_set_elem_ops(handle, eleshp, eledet, elemass, mat, ss, phi, eleporep, stressArr, pstrain) &
                bind(C, name='eqd_set_elem_ops')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle, eleshp, eledet, elemass, mat, ss, phi, eleporep, stressArr, pstrain
        end function
        integer(c_int) function eqd_set_nodal(handle, nodalMassArr, fnms, v1, velArr, dispArr, nodalForceArr) &
                bind(C, name='eqd_set_nodal')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle, nodalMassArr, fnms, v1, velArr, dispArr, nodalForceArr
        end function
        integer(c_int) function eqd_set_fault(handle, nftmx, nftnd, nsmp, un, us, ud, arn, fric, fnft) bind(C, name='eqd_set_fault')
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: nftmx
            type(c_ptr), value :: nftnd, nsmp, un, us, ud, arn, fric, fnft
        end function
        integer(c_int) function eqd_set_halo(handle, numcount, fltnum, fltMPI, fltl, fltr, fltf, fltb, fltd, fltu) &
                bind(C, name='eqd_set_halo')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle, numcount, fltnum, fltMPI, fltl, fltr, fltf, fltb, fltd, fltu
        end function
        integer(c_int) function eqd_set_stations(handle, idhist, nOff, anonfs, nOn, surfaceNodeIdArr, nSurf) &
                bind(C, name='eqd_set_stations')
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value :: handle, idhist, anonfs, surfaceNodeIdArr
            integer(c_int32_t), value :: nOff, nOn, nSurf
        end function
        integer(c_int) function eqd_get_unique_id(id128) bind(C, name='eqd_get_unique_id')
            import :: c_int, c_char
            character(kind=c_char) :: id128(128)
        end function
        integer(c_int) function eqd_set_comm(handle, id128, nranks, rank) bind(C, name='eqd_set_comm')
            import :: c_int, c_ptr, c_char, c_int32_t
            type(c_ptr), value :: handle
            character(kind=c_char) :: id128(128)
            integer(c_int32_t), value :: nranks, rank
        end function
        ! fn is a bind(C) function  integer(c_int32_t) fn(ctx, send, bytes, recv)  that wraps
        ! MPI_Allgather(send, bytes, MPI_BYTE, recv, bytes, MPI_BYTE, MPI_COMM_WORLD); pass it as c_funloc(fn)
        integer(c_int) function eqd_set_host_comm(handle, nranks, rank, fn, ctx) bind(C, name='eqd_set_host_comm')
            import :: c_int, c_ptr, c_funptr, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: nranks, rank
            type(c_funptr), value :: fn
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function eqd_sum_shared(handle) bind(C, name='eqd_sum_shared')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
        end function
        integer(c_int) function eqd_run(handle, nt_begin, nt_end) bind(C, name='eqd_run')
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: nt_begin, nt_end
        end function
        integer(c_int) function eqd_run_group(handles, n, nt_begin, nt_end) bind(C, name='eqd_run_group')
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr) :: handles(*)
            integer(c_int32_t), value :: n, nt_begin, nt_end
        end function
        integer(c_int) function eqd_fetch(handle, which, dst, dst_bytes) bind(C, name='eqd_fetch')
            import :: c_int, c_ptr, c_int32_t, c_int64_t
            type(c_ptr), value :: handle, dst
            integer(c_int32_t), value :: which
            integer(c_int64_t), value :: dst_bytes
        end function
        integer(c_int) function eqd_get_counts(handle, n_regular, n_pml, n_pairs, launches) bind(C, name='eqd_get_counts')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), intent(out) :: n_regular, n_pml, n_pairs, launches
        end function
        integer(c_int) function eqd_get_box_counts(handle, n_regular_box, n_pml_box) bind(C, name='eqd_get_box_counts')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), intent(out) :: n_regular_box, n_pml_box
        end function
        integer(c_int) function eqd_get_timing(handle, ms_slots) bind(C, name='eqd_get_timing')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(out) :: ms_slots(10)
        end function
        integer(c_int) function eqd_get_halo_mode(handle) bind(C, name='eqd_get_halo_mode')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
        end function
        integer(c_int) function eqd_get_march_counts(handle, out5) bind(C, name='eqd_get_march_counts')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), intent(out) :: out5(8)
        end function
        integer(c_int) function eqd_set_option(handle, key, val) bind(C, name='eqd_set_option')
            import :: c_int, c_ptr, c_char, c_int32_t
            type(c_ptr), value :: handle
            character(kind=c_char) :: key(*)
            integer(c_int32_t), value :: val
        end function
        integer(c_int) function eqd_compute_elem_ops(handle, mat, eleporep, stressArr, pstrain) &
                bind(C, name='eqd_compute_elem_ops')
            import :: c_int, c_ptr
            type(c_ptr), value :: handle, mat, eleporep, stressArr, pstrain
        end function
        integer(c_int) function eqd_plan_check(Nn, Ne, nodeElemIdRelation, elemTypeArr, numOfDofPerNodeArr, stats) &
                bind(C, name='eqd_plan_check')
            import :: c_int, c_int32_t, c_int64_t
            integer(c_int32_t), value :: Nn, Ne
            integer(c_int32_t) :: nodeElemIdRelation(8,*), elemTypeArr(*), numOfDofPerNodeArr(*)
            integer(c_int64_t), intent(out) :: stats(24)
        end function
        integer(c_int) function eqd_plan_bank_model(Nn, Ne, nodeElemIdRelation, elemTypeArr, numOfDofPerNodeArr, bank_order, wavefronts) &
                bind(C, name='eqd_plan_bank_model')
            import :: c_int, c_int32_t, c_int64_t
            integer(c_int32_t), value :: Nn, Ne, bank_order
            integer(c_int32_t) :: nodeElemIdRelation(8,*), elemTypeArr(*), numOfDofPerNodeArr(*)
            integer(c_int64_t), intent(out) :: wavefronts(9)
        end function
        integer(c_int) function eqd_box_check(Nn, Ne, meshCoor, nodeElemIdRelation, elemTypeArr, eleshp, phi, ss, nBox, dev) &
                bind(C, name='eqd_box_check')
            import :: c_int, c_int32_t, c_int64_t, c_double
            integer(c_int32_t), value :: Nn, Ne
            real(c_double) :: meshCoor(3,*), eleshp(3,8,*), phi(8,4,*), ss(6,*)
            integer(c_int32_t) :: nodeElemIdRelation(8,*), elemTypeArr(*)
            integer(c_int64_t), intent(out) :: nBox
            real(c_double), intent(out) :: dev(3)
        end function
        integer(c_int) function eqd_march_emulate(Nn, Ne, meshCoor, nodeElemIdRelation, elemTypeArr, numOfDofPerNodeArr, grid, &
                eleshp, ss, eledet, mat, stress6, vel, disp, mass, dt, rdampk, w, update, fsum, fusedFlag, inBundle, stats) &
                bind(C, name='eqd_march_emulate')
            import :: c_int, c_int32_t, c_int64_t, c_double
            integer(c_int32_t), value :: Nn, Ne, grid, update
            real(c_double) :: meshCoor(3,*), eleshp(3,8,*), ss(6,*), eledet(*), mat(*), stress6(6,*), vel(3,*), disp(3,*), mass(*), fsum(3,*)
            integer(c_int32_t) :: nodeElemIdRelation(8,*), elemTypeArr(*), numOfDofPerNodeArr(*), fusedFlag(*), inBundle(*)
            real(c_double), value :: dt, rdampk, w
            integer(c_int64_t), intent(out) :: stats(8)
        end function
        integer(c_int) function eqd_march_pml_emulate(Nn, Ne, meshCoor, nodeElemIdRelation, elemTypeArr, numOfDofPerNodeArr, grid, &
                eleshp, ss, eledet, mat, damps, stress21, vel, disp, dt, rdampk, w, f12, inBundle, stats) &
                bind(C, name='eqd_march_pml_emulate')
            import :: c_int, c_int32_t, c_int64_t, c_double
            integer(c_int32_t), value :: Nn, Ne, grid
            real(c_double) :: meshCoor(3,*), eleshp(3,8,*), ss(6,*), eledet(*), mat(*), damps(3,*), stress21(21,*), vel(3,*), disp(3,*), f12(12,*)
            integer(c_int32_t) :: nodeElemIdRelation(8,*), elemTypeArr(*), numOfDofPerNodeArr(*), inBundle(*)
            real(c_double), value :: dt, rdampk, w
            integer(c_int64_t), intent(out) :: stats(8)
        end function
    end interface

contains

    ! the eqd_allgather_fn of eqd_set_host_comm over the reference's communicator (eqdyna3d.f90:19-21)
    integer(c_int32_t) function eqd_mpi_allgather(ctx, send, nbytes, recv) bind(C)
        include 'mpif.h'
        type(c_ptr), value :: ctx, send, recv
        integer(c_int64_t), value :: nbytes
        character(kind=c_char), pointer :: s(:), r(:)
        integer :: nranks, mpierr
        call MPI_Comm_size(MPI_COMM_WORLD, nranks, mpierr)
        call c_f_pointer(send, s, [nbytes])
        call c_f_pointer(recv, r, [nbytes*nranks])
        call MPI_Allgather(s, int(nbytes), MPI_BYTE, r, int(nbytes), MPI_BYTE, MPI_COMM_WORLD, mpierr)
        eqd_mpi_allgather = mpierr
    end function eqd_mpi_allgather
end module eqdyna_cuda_iface
