! Replacement for src/driver.f90: the explicit time-stepping loop runs on the GPU.
! Same entry point, same pre-state (everything eqdyna3d.f90:33-70 built), same
! post-state (what library_output.f90 reads at eqdyna3d.f90:75-79).
!
! Not compiled in the development image (no gfortran).  Requires the TARGET attribute
! on the globalvar arrays passed through c_loc (src/globalvar.f90:81-101).
subroutine driver

    use globalvar
    use eqdyna_cuda_iface
    use, intrinsic :: iso_c_binding
    implicit none
    include 'mpif.h'

    type(eqd_params) :: p
    type(c_ptr) :: h
    integer(c_int) :: ierr
    integer :: mpierr, nSurf, chunk, nt0, nt1
    character(kind=c_char) :: id128(128), msg(512)
    integer(c_int32_t), target :: fltMPIint(6)

    ! ---- scalars of module globalvar (readInputFiles.f90:182-186 already applied rdampk = rdampk*dt)
    p%dt = dt; p%nstep = nstep; p%me = me; p%npx = npx; p%npy = npy; p%npz = npz
    p%rdampk = rdampk; p%rdampm = rdampm; p%w = w; p%grav = grav; p%roumax = roumax; p%rhow = rhow; p%gamar = gamar
    p%ccosphi = ccosphi; p%sinphi = sinphi; p%tv = tv; p%kapa_hg = kapa_hg; p%dx = dx
    p%C_elastic = C_elastic; p%C_Q = C_Q; p%C_hg = C_hg
    p%PMLb = PMLb; p%nPML = nPML; p%R = R; p%vmaxPML = vmaxPML
    p%friclaw = friclaw; p%C_nuclea = C_nuclea; p%nucfault = nucfault; p%TPV = TPV
    p%insertFaultType = insertFaultType; p%ntotft = ntotft
    p%nucR = nucR; p%nucT = nucT; p%nucRuptVel = nucRuptVel; p%nucdtau0 = nucdtau0
    p%xsource = xsource; p%ysource = ysource; p%zsource = zsource
    p%slipRateThres = slipRateThres; p%tol = tol; p%fric_tp_h = fric_tp_h
    p%outputGroundMotion = outputGroundMotion
    p%reserved_i = 0; p%reserved_d = 0.0d0

    ierr = eqd_create(p, -1_c_int, h)
    if (ierr /= EQD_OK) stop 3

    ierr = eqd_set_mesh(h, totalNumOfNodes, totalNumOfElements, totalNumOfEquations, sizeOfEqNumIndexArr, &
        c_loc(meshCoor), c_loc(nodeElemIdRelation), c_loc(elemTypeArr), c_loc(numOfDofPerNodeArr), &
        c_loc(eqNumStartIndexLoc), c_loc(eqNumIndexArr), c_loc(stressCompIndexArr), 5*sizeOfEqNumIndexArr)
    call check('eqd_set_mesh')
    ierr = eqd_set_elem_ops(h, c_loc(eleshp), c_loc(eledet), c_loc(elemass), c_loc(mat), c_loc(ss), c_loc(phi), &
        c_loc(eleporep), c_loc(stressArr), c_loc(pstrain))
    call check('eqd_set_elem_ops')
    ! mass and fnms were already summed over rank faces by assembleGlobalMass (MPI4NodalQuant), arn by MPI4arn
    ierr = eqd_set_nodal(h, c_loc(nodalMassArr), c_loc(fnms), c_loc(v1), c_loc(velArr), c_loc(dispArr), c_loc(nodalForceArr))
    call check('eqd_set_nodal')
    ierr = eqd_set_fault(h, nftmx, c_loc(nftnd), c_loc(nsmp), c_loc(un), c_loc(us), c_loc(ud), c_loc(arn), c_loc(fric), c_loc(fnft))
    call check('eqd_set_fault')
    fltMPIint = 0
    where (fltMPI) fltMPIint = 1
    ierr = eqd_set_halo(h, c_loc(numcount), c_loc(fltnum), c_loc(fltMPIint), ptr_or_null(fltl, fltnum(1)), &
        ptr_or_null(fltr, fltnum(2)), ptr_or_null(fltf, fltnum(3)), ptr_or_null(fltb, fltnum(4)), &
        ptr_or_null(fltd, fltnum(5)), ptr_or_null(fltu, fltnum(6)))
    call check('eqd_set_halo')
    nSurf = surface_nnode
    ierr = eqd_set_stations(h, c_loc(idhist), numOfOffFaultStCount, c_loc(anonfs), numOfOnFaultStCount, &
        c_loc(surfaceNodeIdArr), nSurf)
    call check('eqd_set_stations')

    ! Optional fast path (DESIGN.md 3d): closed-form operators on tiles of axis-aligned hexahedra.  Off by
    ! default, so that the stored eleshp / phi / ss are used for every element; results with it agree with
    ! the default to ~1e-14 per step (parity tests: 1e-10 against the CPU oracle).
    ! ierr = eqd_set_option(h, 'box'//c_null_char, 2_c_int32_t);          call check('eqd_set_option box')
    ! ierr = eqd_set_option(h, 'box_compact'//c_null_char, 1_c_int32_t);  call check('eqd_set_option box_compact')

    if (npx*npy*npz > 1) then
        if (me == masterProcsId) ierr = eqd_get_unique_id(id128)
        call MPI_Bcast(id128, 128, MPI_CHARACTER, masterProcsId, MPI_COMM_WORLD, mpierr)
        ierr = eqd_set_comm(h, id128, npx*npy*npz, me)
        call check('eqd_set_comm')
    endif

    ! ---- the loop of driver.f90:9-34, in chunks of 100 steps so that the banner keeps appearing
    chunk = 100
    nt0 = 1
    do while (nt0 <= nstep)
        nt1 = min(nt0 + chunk - 1, nstep)
        if (me == masterProcsId) then
            write(*,*) '=     Current time in dynamic rupture                               ='
            write(*,'(X,A,40X,f7.3,4X,A)') '=',  (nt0-1)*dt + dt , 's'
        endif
        ierr = eqd_run(h, nt0, nt1)
        call check('eqd_run')
        nt0 = nt1 + 1
    enddo
    nt = nstep
    timeElapsed = nstep*dt

    ! ---- post-state for output_onfault_st / output_offfault_st / output_frt / output_plastic_strain
    ierr = eqd_fetch(h, EQD_F_FRIC, c_loc(fric), int(8*size(fric), c_int64_t));                 call check('fetch fric')
    ierr = eqd_fetch(h, EQD_F_FNFT, c_loc(fnft), int(8*size(fnft), c_int64_t));                 call check('fetch fnft')
    ierr = eqd_fetch(h, EQD_F_ONFAULT_HIST, c_loc(onFaultQuantHistSCECForm), int(8*size(onFaultQuantHistSCECForm), c_int64_t))
    call check('fetch on-fault stations')
    if (numOfOffFaultStCount > 0) then
        ierr = eqd_fetch(h, EQD_F_OFFFAULT_HIST, c_loc(OffFaultStGramSCEC), int(8*size(OffFaultStGramSCEC), c_int64_t))
        call check('fetch off-fault stations')
    endif
    ierr = eqd_fetch(h, EQD_F_DISP, c_loc(dispArr), int(8*size(dispArr), c_int64_t));           call check('fetch disp')
    ierr = eqd_fetch(h, EQD_F_VEL, c_loc(velArr), int(8*size(velArr), c_int64_t));              call check('fetch vel')
    if (C_elastic == 0) then
        ierr = eqd_fetch(h, EQD_F_PSTRAIN, c_loc(pstrain), int(8*size(pstrain), c_int64_t));    call check('fetch pstrain')
    endif
    ierr = eqd_destroy(h)

contains

    subroutine check(what)
        character(len=*), intent(in) :: what
        integer(c_int) :: rc2
        if (ierr == EQD_OK) return
        rc2 = eqd_last_error(h, msg, 512_c_int)
        write(*,*) 'eqdyna_b200: ', what, ' failed with code ', ierr
        write(*,*) msg
        ! mirrors the reference's stop sites: 1 NaN velocity (driver.f90:147-152), 2 negative damping (comdampv.f90:114-118)
        call MPI_Abort(MPI_COMM_WORLD, ierr, mpierr)
        stop
    end subroutine check

    function ptr_or_null(a, n) result(p0)
        integer(kind=4), allocatable, target, intent(in) :: a(:)
        integer(kind=4), intent(in) :: n
        type(c_ptr) :: p0
        p0 = c_null_ptr
        if (n > 0 .and. allocated(a)) p0 = c_loc(a)
    end function ptr_or_null

end subroutine driver
