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
    integer :: mpierr, nSurf, chunk, nt0, nt1, k, i, nGmDone, nGmNow, hypoPair
    character(kind=c_char) :: id128(128), msg(512)
    integer(c_int32_t), target :: fltMPIint(6)
    real(c_double), allocatable, target :: hypoLog(:,:), gmBuf(:,:,:), srcBuf(:,:)

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
    ! The marching kernel (DESIGN.md 3e; elastic runs) is chosen BEFORE eqd_set_mesh:
    ! ierr = eqd_set_option(h, 'march'//c_null_char, 1_c_int32_t)

    ! one box (<= 8 GPUs, peers mappable): the library's set-up exchanges go through MPI_Allgather and the steps over
    ! peer memory.  Across boxes (or with EQD_USE_NCCL set) give it an NCCL communicator as well: the step exchange
    ! then falls back to ncclSend / ncclRecv where a neighbour's memory cannot be mapped.
    if (npx*npy*npz > 1) then
        ierr = eqd_set_host_comm(h, npx*npy*npz, me, c_funloc(eqd_mpi_allgather), c_null_ptr)
        call check('eqd_set_host_comm')
        call get_environment_variable('EQD_USE_NCCL', length=k)
        if (k > 0) then
            if (me == masterProcsId) ierr = eqd_get_unique_id(id128)
            call MPI_Bcast(id128, 128, MPI_CHARACTER, masterProcsId, MPI_COMM_WORLD, mpierr)
            ierr = eqd_set_comm(h, id128, npx*npy*npz, me)
            call check('eqd_set_comm')
        endif
    endif

    ! ---- the loop of driver.f90:9-34, in chunks of 100 steps: the banner of driver.f90:13-17 appears at
    ! nt = 1, 101, ... with the accumulated timeElapsed, and after every chunk the samples output_gm /
    ! output_src_evol would have appended inside it (driver.f90:30-33: every step with mod(nt,10) == 1) are
    ! fetched and appended to gm<me> / src_evol<me>, and the hypocentre block of showSourceDynamics
    ! (faulting.f90:343-365) is printed from the recorded log -- one block per step, as the reference does.
    allocate(hypoLog(13, nstep))
    hypoPair = 0
    if (nftnd(1) > 0) then
        do i = 1, nftnd(1)
            if (abs(meshCoor(1,nsmp(1,i,1))-xsource) < tol .and. abs(meshCoor(3,nsmp(1,i,1))-zsource) < tol) hypoPair = i
        enddo
    endif
    chunk = 100
    nt0 = 1
    nGmDone = 0
    do while (nt0 <= nstep)
        nt1 = min(nt0 + chunk - 1, nstep)
        if (me == masterProcsId) then
            write(*,*) '=                                                                   ='
            write(*,*) '=     Current time in dynamic rupture                               ='
            write(*,'(X,A,40X,f7.3,4X,A)') '=',  timeElapsed + dt , 's'
        endif
        ierr = eqd_run(h, nt0, nt1)
        call check('eqd_run')
        do k = nt0, nt1
            timeElapsed = timeElapsed + dt           ! the same accumulated sum as driver.f90:11
        enddo
        nt = nt1
        if (hypoPair > 0) then
            ierr = eqd_fetch(h, EQD_F_HYPO_LOG, c_loc(hypoLog), int(8*size(hypoLog), c_int64_t)); call check('fetch hypocentre log')
            do k = nt0, nt1
                write(*,*) 'Hypocenter dynamics'
                write(*,'(X,A,3E15.7)') 'X, Y, Z      (m)   ', meshCoor(1,nsmp(1,hypoPair,1)), meshCoor(2,nsmp(1,hypoPair,1)), meshCoor(3,nsmp(1,hypoPair,1))
                write(*,'(X,A,E15.7)') 'TimeElapsed  (s)   ', hypoLog(1,k)
                write(*,'(X,A,3E15.7)') 'n,s,d tract  (MPa) ', hypoLog(2,k)/1.d6, hypoLog(3,k)/1.d6, hypoLog(4,k)/1.d6
                write(*,'(X,A,E15.7)') 'state_normal (MPa) ', hypoLog(5,k)/1.d6
                write(*,'(X,A,3E15.7)') 'n,s,d slip   (m)   ', hypoLog(6,k), hypoLog(7,k), hypoLog(8,k)
                write(*,'(X,A,3E15.7)') 's,d, peak sr (m/s) ', hypoLog(9,k), hypoLog(10,k), hypoLog(11,k)
                write(*,'(X,A,E15.7)') 'cummul slip  (m)   ', hypoLog(12,k)
                write(*,'(X,A,3E15.7)') 'sw_fs, sw_fd, sw_D0', fric(1,hypoPair,1), fric(2,hypoPair,1), fric(3,hypoPair,1)
                write(*,'(X,A,2E15.7)') 'rsf_a, rsf_b       ', fric(9,hypoPair,1), fric(10,hypoPair,1)
                write(*,'(X,A,E15.7)') 'rsf_state          ', hypoLog(13,k)
                write(*,'(X,A,E15.7)') 'Nuc add tau0 (MPa) ', fric(81,hypoPair,1)/1.d6
            enddo
        endif
        if (outputGroundMotion == 1) then
            nGmNow = (nt1 - 1)/10 + 1                 ! samples taken so far: steps 1, 11, 21, ... <= nt1
            if (nGmNow > nGmDone) then
                if (surface_nnode > 0) then            ! output_gm, library_output.f90:267-279
                    allocate(gmBuf(3, surface_nnode, nGmNow))
                    ierr = eqd_fetch(h, EQD_F_GM, c_loc(gmBuf), int(8*size(gmBuf), c_int64_t)); call check('fetch gm')
                    open(unit=10009+me, file='gm'//mm, status='unknown', position='append', access='stream')
                    write(10009+me) gmBuf(:, :, nGmDone+1:nGmNow)
                    close(10009+me)
                    deallocate(gmBuf)
                endif
                if (nftnd(1) > 0) then                 ! output_src_evol, library_output.f90:297-312
                    allocate(srcBuf(nftnd(1), nGmNow))
                    ierr = eqd_fetch(h, EQD_F_SRC_EVOL, c_loc(srcBuf), int(8*size(srcBuf), c_int64_t)); call check('fetch src_evol')
                    open(unit=30009+me, file='src_evol'//mm, position='append', access='stream')
                    write(30009+me) srcBuf(:, nGmDone+1:nGmNow)
                    close(30009+me)
                    deallocate(srcBuf)
                endif
                nGmDone = nGmNow
            endif
        endif
        nt0 = nt1 + 1
    enddo
    nt = nstep
    deallocate(hypoLog)

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
    ! output_plastic_strain (library_output.f90:221-244) prints stressArr next to pstrain: the final stresses
    ierr = eqd_fetch(h, EQD_F_STRESS, c_loc(stressArr), int(8*size(stressArr), c_int64_t));     call check('fetch stress')
    ! v1 and the accelerations the reference leaves in nodalForceArr (driver.f90:29): needed by a host that goes on
    ierr = eqd_fetch(h, EQD_F_V1, c_loc(v1), int(8*size(v1), c_int64_t));                       call check('fetch v1')
    ierr = eqd_fetch(h, EQD_F_FORCE, c_loc(nodalForceArr), int(8*size(nodalForceArr), c_int64_t)); call check('fetch accel')
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
