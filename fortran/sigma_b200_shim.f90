!==========================================================================!
!==========================================================================!
module sigma_b200_shim                                                     !
!==========================================================================!
!==========================================================================!
!==== iso_c_binding interface to libsigma_b200.so (include/sigma_b200.h) ==!
!==== and the marshalling helpers the SiGMA procedure bodies call.      ====!
!====                                                                  ====!
!==== Not compiled in this repository's image (it has no Fortran        ====!
!==== compiler); every behavioural decision lives on the C side, where  ====!
!==== it is tested.  tests/cxx/*.cpp make exactly the calls below from   ====!
!==== C++ (sigma_b200/host/sigma.hpp) and are run on the GPU.            ====!
!====                                                                  ====!
!==== INTEGRATION.md shows, procedure by procedure, which reference     ====!
!==== body is replaced by which call of this module.                    ====!
!==========================================================================!
!==========================================================================!

use iso_c_binding

implicit none

integer(c_int), parameter :: SIGB_OK = 0, SIGB_ROW = 0, SIGB_COL = 1

! number of GPUs of the single-process multi-GPU mode (0: one GPU, the default);
! set by sigma_use_gpus below
integer(c_int), save :: sigma_gpus_in_use = 0


!--------------------------------------------------------------------------!
interface                                                                  !
!--------------------------------------------------------------------------!
    ! runtime
    function sigb_init(device) bind(c, name='sigb_init') result(stat)
        import :: c_int
        integer(c_int), value :: device
        integer(c_int) :: stat
    end function

    function sigb_last_error() bind(c, name='sigb_last_error') result(msg)
        import :: c_ptr
        type(c_ptr) :: msg
    end function

    ! graphs
    function sigb_cs_graph_create(n, m, ptr1, node1, order, g) &
            & bind(c, name='sigb_cs_graph_create') result(stat)
        import :: c_int, c_int32_t, c_ptr
        integer(c_int32_t), value :: n, m
        integer(c_int32_t), intent(in) :: ptr1(*), node1(*)
        integer(c_int), value :: order
        type(c_ptr), intent(out) :: g
        integer(c_int) :: stat
    end function

    function sigb_ell_graph_create(n, m, max_d, node_cm, degrees, g) &
            & bind(c, name='sigb_ell_graph_create') result(stat)
        import :: c_int, c_int32_t, c_ptr
        integer(c_int32_t), value :: n, m, max_d
        integer(c_int32_t), intent(in) :: node_cm(*), degrees(*)
        type(c_ptr), intent(out) :: g
        integer(c_int) :: stat
    end function

    function sigb_graph_retain(g) bind(c, name='sigb_graph_retain') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: g
        integer(c_int) :: stat
    end function

    function sigb_graph_release(g) bind(c, name='sigb_graph_release') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: g
        integer(c_int) :: stat
    end function

    ! matrices
    function sigb_matrix_create(g, A) bind(c, name='sigb_matrix_create') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: g
        type(c_ptr), intent(out) :: A
        integer(c_int) :: stat
    end function

    function sigb_matrix_set_values(A, val, count) &
            & bind(c, name='sigb_matrix_set_values') result(stat)
        import :: c_int, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: A
        real(c_double), intent(in) :: val(*)
        integer(c_int64_t), value :: count
        integer(c_int) :: stat
    end function

    function sigb_matrix_destroy(A) bind(c, name='sigb_matrix_destroy') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: A
        integer(c_int) :: stat
    end function

    ! matvec
    function sigb_matvec(A, trans, x, y) bind(c, name='sigb_matvec') result(stat)
        import :: c_int, c_double, c_ptr
        type(c_ptr), value :: A
        integer(c_int), value :: trans
        real(c_double), intent(in) :: x(*)
        real(c_double), intent(out) :: y(*)
        integer(c_int) :: stat
    end function

    function sigb_matvec_add(A, trans, x, y) bind(c, name='sigb_matvec_add') result(stat)
        import :: c_int, c_double, c_ptr
        type(c_ptr), value :: A
        integer(c_int), value :: trans
        real(c_double), intent(in) :: x(*)
        real(c_double), intent(inout) :: y(*)
        integer(c_int) :: stat
    end function

    ! solvers
    function sigb_cg_create(tolerance, s) bind(c, name='sigb_cg_create') result(stat)
        import :: c_int, c_double, c_ptr
        real(c_double), value :: tolerance
        type(c_ptr), intent(out) :: s
        integer(c_int) :: stat
    end function

    function sigb_bicgstab_create(tolerance, s) bind(c, name='sigb_bicgstab_create') result(stat)
        import :: c_int, c_double, c_ptr
        real(c_double), value :: tolerance
        type(c_ptr), intent(out) :: s
        integer(c_int) :: stat
    end function

    function sigb_jacobi_create(s) bind(c, name='sigb_jacobi_create') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), intent(out) :: s
        integer(c_int) :: stat
    end function

    function sigb_ldu_create(s) bind(c, name='sigb_ldu_create') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), intent(out) :: s
        integer(c_int) :: stat
    end function

    function sigb_ldu_get_sizes(s, n, nL, nU, nflev, nblev) &
            & bind(c, name='sigb_ldu_get_sizes') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_ptr
        type(c_ptr), value :: s
        integer(c_int32_t), intent(out) :: n, nflev, nblev
        integer(c_int64_t), intent(out) :: nL, nU
        integer(c_int) :: stat
    end function

    function sigb_ldu_get_factors(s, Lptr, Lnode, Lval, Uptr, Unode, Uval, D) &
            & bind(c, name='sigb_ldu_get_factors') result(stat)
        import :: c_int, c_int32_t, c_double, c_ptr
        type(c_ptr), value :: s
        integer(c_int32_t), intent(out) :: Lptr(*), Lnode(*), Uptr(*), Unode(*)
        real(c_double), intent(out) :: Lval(*), Uval(*), D(*)
        integer(c_int) :: stat
    end function

    function sigb_solver_setup(s, A) bind(c, name='sigb_solver_setup') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: s, A
        integer(c_int) :: stat
    end function

    function sigb_solver_set_params(s, tolerance) &
            & bind(c, name='sigb_solver_set_params') result(stat)
        import :: c_int, c_double, c_ptr
        type(c_ptr), value :: s
        real(c_double), value :: tolerance
        integer(c_int) :: stat
    end function

    function sigb_solver_solve(s, A, x, b, pc) bind(c, name='sigb_solver_solve') result(stat)
        import :: c_int, c_double, c_ptr
        type(c_ptr), value :: s, A, pc          ! pc = c_null_ptr: unpreconditioned
        real(c_double), intent(inout) :: x(*)
        real(c_double), intent(in) :: b(*)
        integer(c_int) :: stat
    end function

    function sigb_solver_get_info(s, iterations, res2, capped) &
            & bind(c, name='sigb_solver_get_info') result(stat)
        import :: c_int, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: s
        integer(c_int64_t), intent(out) :: iterations
        real(c_double), intent(out) :: res2
        integer(c_int), intent(out) :: capped
        integer(c_int) :: stat
    end function

    function sigb_solver_destroy(s) bind(c, name='sigb_solver_destroy') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: s
        integer(c_int) :: stat
    end function

    ! eigensolver
    function sigb_lanczos(A, n, q1, seed, T, Q) bind(c, name='sigb_lanczos') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: A
        integer(c_int32_t), value :: n
        type(c_ptr), value :: q1                ! c_null_ptr: library draws the start vector
        integer(c_int64_t), value :: seed
        real(c_double), intent(out) :: T(3, *), Q(*)
        integer(c_int) :: stat
    end function

    function sigb_eigensolve(A, n, q1, seed, lambda, V) &
            & bind(c, name='sigb_eigensolve') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: A
        integer(c_int32_t), value :: n
        type(c_ptr), value :: q1
        integer(c_int64_t), value :: seed
        real(c_double), intent(out) :: lambda(*), V(*)
        integer(c_int) :: stat
    end function

    function sigb_generalized_lanczos(A, B, b_solver, b_pc, n, q1, seed, T, Q) &
            & bind(c, name='sigb_generalized_lanczos') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: A, B, b_solver, b_pc   ! b_pc = c_null_ptr: none attached
        integer(c_int32_t), value :: n
        type(c_ptr), value :: q1
        integer(c_int64_t), value :: seed
        real(c_double), intent(out) :: T(3, *), Q(*)
        integer(c_int) :: stat
    end function

    function sigb_generalized_eigensolve(A, B, b_solver, b_pc, n, q1, seed, lambda, V) &
            & bind(c, name='sigb_generalized_eigensolve') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: A, B, b_solver, b_pc   ! b_pc = c_null_ptr: none attached
        integer(c_int32_t), value :: n
        type(c_ptr), value :: q1
        integer(c_int64_t), value :: seed
        real(c_double), intent(out) :: lambda(*), V(*)
        integer(c_int) :: stat
    end function

    ! operator expressions: every constructor returns another operator handle
    ! that sigb_matvec*, sigb_solver_* and sigb_lanczos* accept unchanged
    function sigb_operator_sum(A, B, C) bind(c, name='sigb_operator_sum') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: A, B
        type(c_ptr), intent(out) :: C
        integer(c_int) :: stat
    end function

    function sigb_operator_product(A, B, C) bind(c, name='sigb_operator_product') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: A, B
        type(c_ptr), intent(out) :: C
        integer(c_int) :: stat
    end function

    function sigb_operator_adjoint(A, B) bind(c, name='sigb_operator_adjoint') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: A
        type(c_ptr), intent(out) :: B
        integer(c_int) :: stat
    end function

    function sigb_composite_create(num_row_mats, num_col_mats, rows, cols, blocks, A) &
            & bind(c, name='sigb_composite_create') result(stat)
        import :: c_int, c_int32_t, c_ptr
        integer(c_int32_t), value :: num_row_mats, num_col_mats
        integer(c_int32_t), intent(in) :: rows(*), cols(*)
        type(c_ptr), intent(in) :: blocks(*)     ! sub_mats(it, jt) at (it-1)*num_col_mats + jt
        type(c_ptr), intent(out) :: A
        integer(c_int) :: stat
    end function

    ! matrix copy / format conversion on the device
    function sigb_matrix_copy(A, frmt, trans, B) bind(c, name='sigb_matrix_copy') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: A
        integer(c_int), value :: frmt, trans     ! frmt: 1 csr, 2 csc, 3 ellpack
        type(c_ptr), intent(out) :: B
        integer(c_int) :: stat
    end function

    function sigb_matrix_get_format(A, frmt, n_lines, n_ids, ne, max_d) &
            & bind(c, name='sigb_matrix_get_format') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_ptr
        type(c_ptr), value :: A
        integer(c_int), intent(out) :: frmt
        integer(c_int32_t), intent(out) :: n_lines, n_ids, max_d
        integer(c_int64_t), intent(out) :: ne
        integer(c_int) :: stat
    end function

    function sigb_matrix_get_arrays(A, ptr_or_degrees, node, val) &
            & bind(c, name='sigb_matrix_get_arrays') result(stat)
        import :: c_int, c_int32_t, c_double, c_ptr
        type(c_ptr), value :: A
        integer(c_int32_t), intent(out) :: ptr_or_degrees(*), node(*)
        real(c_double), intent(out) :: val(*)
        integer(c_int) :: stat
    end function

    function sigb_matrix_add_values(A, count, i1, j1, z) &
            & bind(c, name='sigb_matrix_add_values') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: A
        integer(c_int64_t), value :: count
        integer(c_int32_t), intent(in) :: i1(*), j1(*)
        real(c_double), intent(in) :: z(*)
        integer(c_int) :: stat
    end function

    function sigb_matrix_retain(A) bind(c, name='sigb_matrix_retain') result(stat)
        import :: c_int, c_ptr
        type(c_ptr), value :: A
        integer(c_int) :: stat
    end function

    ! single-process multi-GPU mode: the whole pattern in, one row block per GPU behind one handle;
    ! every other entry point above then takes that handle with the caller's whole arrays
    function sigb_mgpu_init(ndev) bind(c, name='sigb_mgpu_init') result(stat)
        import :: c_int
        integer(c_int), value :: ndev
        integer(c_int) :: stat
    end function

    function sigb_mgpu_finalize() bind(c, name='sigb_mgpu_finalize') result(stat)
        import :: c_int
        integer(c_int) :: stat
    end function

    function sigb_mgpu_device_count(ndev) bind(c, name='sigb_mgpu_device_count') result(stat)
        import :: c_int
        integer(c_int), intent(out) :: ndev
        integer(c_int) :: stat
    end function

    function sigb_mgpu_csr_create(n, ptr1, node1, A) bind(c, name='sigb_mgpu_csr_create') result(stat)
        import :: c_int, c_int32_t, c_ptr
        integer(c_int32_t), value :: n
        integer(c_int32_t), intent(in) :: ptr1(*), node1(*)
        type(c_ptr), intent(out) :: A
        integer(c_int) :: stat
    end function

    ! the graph builders fed by an edge stream (cs_graph_build, ellpack_graph_build), on the device
    function sigb_cs_graph_build(n, m, count, src_i, src_j, trans, order, g) &
            & bind(c, name='sigb_cs_graph_build') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_ptr
        integer(c_int32_t), value :: n, m
        integer(c_int64_t), value :: count
        integer(c_int32_t), intent(in) :: src_i(*), src_j(*)
        integer(c_int), value :: trans, order
        type(c_ptr), intent(out) :: g
        integer(c_int) :: stat
    end function

    function sigb_ell_graph_build(n, m, count, src_i, src_j, trans, g) &
            & bind(c, name='sigb_ell_graph_build') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_ptr
        integer(c_int32_t), value :: n, m
        integer(c_int64_t), value :: count
        integer(c_int32_t), intent(in) :: src_i(*), src_j(*)
        integer(c_int), value :: trans
        type(c_ptr), intent(out) :: g
        integer(c_int) :: stat
    end function

    ! lanczos with Q left on the device (Q_dev, q1_dev are device addresses held as c_ptr)
    function sigb_lanczos_dev(A, n, q1_dev, seed, T, Q_dev) bind(c, name='sigb_lanczos_dev') result(stat)
        import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
        type(c_ptr), value :: A, q1_dev, Q_dev
        integer(c_int32_t), value :: n
        integer(c_int64_t), value :: seed
        real(c_double), intent(out) :: T(*)
        integer(c_int) :: stat
    end function

    function c_strlen(s) bind(c, name='strlen') result(n)
        import :: c_ptr, c_size_t
        type(c_ptr), value :: s
        integer(c_size_t) :: n
    end function
end interface


contains


!--------------------------------------------------------------------------!
subroutine sigma_use_gpus(ndev)                                            !
!--------------------------------------------------------------------------!
! Switch a serial `use sigma` program to all GPUs of the box (ndev <= 0) or  !
! to ndev of them: csr matrices mirrored from now on are row-sharded by the  !
! library (sigb_mgpu_csr_create); nothing else in the program changes.       !
!--------------------------------------------------------------------------!
    integer, intent(in) :: ndev
    call sigb_check( sigb_mgpu_init(int(ndev, c_int)) )
    call sigb_check( sigb_mgpu_device_count(sigma_gpus_in_use) )
end subroutine sigma_use_gpus



!--------------------------------------------------------------------------!
subroutine sigb_check(stat)                                                !
!--------------------------------------------------------------------------!
! The reference's error convention: print a message and `call exit(1)`     !
! (e.g. src/solver/cg_solvers.f90:61-65).                                   !
!--------------------------------------------------------------------------!
    integer(c_int), intent(in) :: stat
    character(kind=c_char), pointer :: msg(:)
    type(c_ptr) :: cmsg

    if (stat /= SIGB_OK) then
        cmsg = sigb_last_error()
        call c_f_pointer(cmsg, msg, [c_strlen(cmsg)])
        print *, msg
        print *, 'Terminating.'
        call exit(1)
    endif

end subroutine sigb_check


end module sigma_b200_shim



!==========================================================================!
!==== How the reference's procedure BODIES change (signatures do not).  ====!
!==== Shown as comments because they are edits to reference files; see  ====!
!==== INTEGRATION.md for the full list.                                  ====!
!==========================================================================!
!
! --- src/graph/formats/cs_graphs.f90, type cs_graph (:11-60) gains
!         type(c_ptr), private :: mirror = c_null_ptr      ! device pattern
!         logical, private :: mirror_is_col = .false.
!     every mutator (add_edge :400, delete_edge, left/right_permute :499-571,
!     build :109) ends with       call g%drop_mirror()
!     (sigb_graph_release + mirror = c_null_ptr).
!
! --- src/matrix/formats/cs_matrices.f90, type cs_matrix (:32-107) gains
!         type(c_ptr), private :: mirror = c_null_ptr      ! device values
!         logical, private :: dirty = .true.
!     every mutator (set_value/add_value :840-947, zero, scalar_multiply
!     :448-490, permutes :972-1099) sets A%dirty = .true.
!
!     subroutine cs_matvec_add(A, x, y)                     ! :500-508
!         class(cs_matrix), intent(in) :: A
!         real(dp), intent(in)    :: x(:)
!         real(dp), intent(inout) :: y(:)
!         call A%sync_mirror()
!         call sigb_check( sigb_matvec_add(A%mirror, 0_c_int, x, y) )
!     end subroutine
!
!     subroutine cs_matvec_t_add(A, x, y)                   ! :513-521
!         call A%sync_mirror()
!         call sigb_check( sigb_matvec_add(A%mirror, 1_c_int, x, y) )
!     end subroutine
!
!     subroutine sync_mirror(A)                             ! new, private
!         class(cs_matrix), intent(inout) :: A
!         integer(c_int) :: order
!         ! single-process multi-GPU mode (after sigb_mgpu_init): the whole pattern
!         ! goes in, the library makes one row block per GPU; set_values, matvec,
!         ! matvec_add and the solvers below then drive all GPUs with whole arrays
!         if (.not. c_associated(A%mirror) .and. sigma_gpus_in_use > 0 &
!                 & .and. .not. A%get_col_is_fast .and. A%nrow == A%ncol) then
!             call sigb_check( sigb_mgpu_csr_create(A%g%n, A%g%ptr, A%g%node, &
!                                         & A%mirror) )
!             A%dirty = .true.
!         endif
!         if (.not. c_associated(A%mirror) .and. .not. c_associated(A%g%mirror)) then
!             order = SIGB_ROW
!             if (A%get_col_is_fast) order = SIGB_COL       ! csc_matrix
!             call sigb_check( sigb_cs_graph_create(A%g%n, A%g%m, A%g%ptr, &
!                                         & A%g%node, order, A%g%mirror) )
!         endif
!         if (.not. c_associated(A%mirror)) then
!             call sigb_check( sigb_matrix_create(A%g%mirror, A%mirror) )
!             A%dirty = .true.
!         endif
!         if (A%dirty) then
!             call sigb_check( sigb_matrix_set_values(A%mirror, A%val, &
!                                         & size(A%val, kind=c_int64_t)) )
!             A%dirty = .false.
!         endif
!     end subroutine
!
! --- src/linear_operator/linear_operator_interface.f90
!     linear_operator_matvec (:185-194) keeps `y = 0; call A%matvec_add(x, y)`
!     for operators without a mirror; cs_matrix / ellpack_matrix override
!     matvec / matvec_t with sigb_matvec (zero-fill fused on the device).
!
! --- src/matrix/formats/ellpack_matrices.f90: same pattern with
!     sigb_ell_graph_create(g%n, g%m, g%max_d, g%node, g%degrees, g%mirror)
!     and sigb_matrix_set_values(A%mirror, A%val, size(A%val)) -- the Fortran
!     arrays node(max_d, n) / val(max_d, n) are passed as they are.
!
! --- src/solver/cg_solvers.f90, type cg_solver (:10-28) gains
!         type(c_ptr), private :: dev = c_null_ptr
!
!     subroutine cg_setup(solver, A)                        ! :52-90
!         ... unchanged checks (non-square -> print + exit(1)) ...
!         if (.not. c_associated(solver%dev)) &
!             call sigb_check( sigb_cg_create(solver%tolerance, solver%dev) )
!         select type(A)
!             class is (cs_matrix)                          ! or ellpack_matrix
!                 call A%sync_mirror()
!                 call sigb_check( sigb_solver_setup(solver%dev, A%mirror) )
!         end select
!         solver%iterations = 0
!     end subroutine
!
!     subroutine cg_solve(solver, A, x, b)                  ! :116-150
!         integer(c_int64_t) :: it ; real(c_double) :: res2 ; integer(c_int) :: capped
!         call A%sync_mirror()
!         call sigb_check( sigb_solver_solve(solver%dev, A%mirror, x, b, c_null_ptr) )
!         call sigb_check( sigb_solver_get_info(solver%dev, it, res2, capped) )
!         solver%iterations = int(it)       ! accumulates across solves, like :145
!     end subroutine
!
!     subroutine cg_solve_pc(solver, A, x, b, pc)           ! :155-194
!         select type(pc)
!             type is (jacobi_solver)
!                 call sigb_check( sigb_solver_solve(solver%dev, A%mirror, x, b, pc%dev) )
!             class default
!                 ... the reference's host loop, unchanged (matvec through the mirror) ...
!         end select
!     end subroutine
!
!     bicgstab_solvers.f90 (:124-237) and jacobi_solvers.f90 (:37-81) follow
!     the same pattern with sigb_bicgstab_create / sigb_jacobi_create.
!
! --- src/eigensolver.f90
!     subroutine lanczos(A, T, Q)                           ! :27-90
!         call init_seed() ; call random_number(Q(:,1)) ; Q(:,1) = 2 * Q(:,1) - 1
!         call sigb_check( sigb_lanczos(A%mirror, size(T, 2), c_loc(Q(1,1)), &
!                                     & 0_c_int64_t, T, Q) )
!     end subroutine
!     (the library normalises the start vector as :52 does; passing
!      c_null_ptr instead lets it draw the vector from `seed`.)
!
! --- src/linear_operator/linear_operator_interface.f90, type linear_operator
!     (:18-45) gains a deferred-with-default accessor every operator answers:
!         procedure :: device_handle => linear_operator_no_mirror   ! c_null_ptr
!     cs_matrix / ellpack_matrix override it with `call A%sync_mirror();
!     h = A%mirror`.  The solvers' `select type(A)` above then becomes
!         h = A%device_handle()
!         if (c_associated(h)) then ; call sigb_check( sigb_solver_setup(solver%dev, h) ) ...
!     so that ANY operator with a device mirror -- a stored matrix or one of
!     the expressions below -- drives the device-resident loops.
!
! --- src/linear_operator/linear_operator_sums.f90, type operator_sum (:11-20)
!     gains  type(c_ptr), private :: mirror = c_null_ptr
!
!     function operator_sum_device_handle(A) result(h)      ! new
!         hA = A%summands(1)%ap%device_handle()              ! refreshes dirty values
!         hB = A%summands(2)%ap%device_handle()
!         if (.not. c_associated(A%mirror)) &
!             call sigb_check( sigb_operator_sum(hA, hB, A%mirror) )
!         h = A%mirror
!     end function
!
!     subroutine operator_sum_matvec_add(A, x, y)           ! :100-114
!         h = A%device_handle()
!         if (c_associated(h)) then
!             call sigb_check( sigb_matvec_add(h, 0_c_int, x, y) )
!         else
!             ... the reference loop over summands, unchanged ...
!         endif
!     end subroutine
!     (matvec_t_add :117-131 with trans = 1; operator_sum_destroy :136-159
!      additionally calls sigb_matrix_destroy(A%mirror).)
!
!     linear_operator_products.f90 (:39-150) and linear_operator_adjoints.f90
!     (:28-86) follow the same pattern with sigb_operator_product /
!     sigb_operator_adjoint; the host scratch vectors z1, z2 (:15,61) are no
!     longer touched when the product has a device mirror.
!
! --- src/matrix/sparse_matrix_composites.f90, type sparse_matrix (:41-49)
!     gains  type(c_ptr), private :: mirror = c_null_ptr ; set_submatrix
!     (:1031-1065) and set_matrix_type (:267-307) drop it.
!
!     function composite_mat_device_handle(A) result(h)     ! new
!         type(c_ptr) :: blocks(A%num_row_mats * A%num_col_mats)
!         do it = 1, A%num_row_mats ; do jt = 1, A%num_col_mats
!             blocks((it - 1) * A%num_col_mats + jt) = A%sub_mats(it, jt)%mat%device_handle()
!         enddo ; enddo
!         if (.not. c_associated(A%mirror)) call sigb_check( sigb_composite_create( &
!                 & A%num_row_mats, A%num_col_mats, &
!                 & A%row_ptr(2:) - A%row_ptr(:A%num_row_mats), &
!                 & A%col_ptr(2:) - A%col_ptr(:A%num_col_mats), blocks, A%mirror) )
!         h = A%mirror
!     end function
!
!     composite_matvec_add (:1076-1100) / composite_matvec_t_add (:1105-1129)
!     become one sigb_matvec_add call on that handle: the block loops run on
!     the device over offset slices of x and y, no host round trip per block.
!
! --- src/eigensolver.f90, generalized_lanczos (:95-155):
!         call sigb_check( sigb_generalized_lanczos(A%device_handle(), B%device_handle(), &
!                 & B%solver%dev, pc_or_null, size(T, 2), c_loc(Q(1,1)), 0_c_int64_t, T, Q) )
!     where B%solver is what `call B%set_solver(...)` attached (:134 runs it).
!
! --- src/eigensolver.f90, generalized_eigensolve (:189-208):
!         call sigb_check( sigb_generalized_eigensolve(A%device_handle(), B%device_handle(), &
!                 & B%solver%dev, pc_or_null, size(lambda), c_null_ptr, 0_c_int64_t, lambda, V) )
!
! --- src/matrix/formats/cs_matrices.f90, cs_matrix_copy_matrix (:294-322):
!         h = B%device_handle()
!         if (c_associated(h)) then
!             frmt = 1 ; if (A%get_col_is_fast) frmt = 2
!             call sigb_check( sigb_matrix_copy(h, frmt, merge(1, 0, tr), A%mirror) )
!             call sigb_check( sigb_matrix_get_format(A%mirror, frmt, n, m, ne, max_d) )
!             allocate(A%g) ; A%g%n = n ; A%g%m = m ; A%g%ne = ne ; A%g%max_d = max_d
!             allocate(A%g%ptr(n + 1), A%g%node(ne), A%val(ne))
!             call sigb_check( sigb_matrix_get_arrays(A%mirror, A%g%ptr, A%g%node, A%val) )
!             A%dirty = .false. ; A%graph_set = .true.
!         else
!             ... build_graph_from_matrix + copy_matrix_values, unchanged ...
!         endif
!     ellpack_matrix_copy_matrix (ellpack_matrices.f90:169-198) likewise with
!     frmt = 3 and A%g%degrees(n), A%g%node(max_d, n), A%val(max_d, n).
!
! --- src/matrix/formats/cs_matrices.f90, csr_matrix_add_multiple_values (:934-967)
!     (csc and ellpack likewise): when every (is(k), js(l)) is in the pattern,
!         n = size(is) * size(js)
!         ii = [(is(k), l = 1, size(js)), k = 1, size(is))]     ! B(k, l) walked k outer, l inner
!         jj = [(js(l), l = 1, size(js)), k = 1, size(is))]
!         zz = [((B(k, l), l = 1, size(js)), k = 1, size(is))]
!         call A%sync_mirror()
!         call sigb_check( sigb_matrix_add_values(A%mirror, n, ii, jj, zz) )
!         call sigb_check( sigb_matrix_get_arrays(A%mirror, dummy_ptr, dummy_node, A%val) )
!     An assembly loop (examples/fem.f90:43-47) gathers its add_value calls into
!     one such batch per mesh instead of one call per entry.
!
! --- src/solver/ldu_solvers.f90, type sparse_ldu_solver (:35-58) gains
!         type(c_ptr), private :: dev = c_null_ptr
!
!     subroutine sparse_ldu_setup(solver, A)                ! :95-130
!         ... unchanged non-square check (print + exit(1)) ...
!         h = A%device_handle()
!         if (c_associated(h)) then
!             if (.not. c_associated(solver%dev)) call sigb_check( sigb_ldu_create(solver%dev) )
!             call sigb_check( sigb_solver_setup(solver%dev, h) )   ! pattern once, numbers every call
!             solver%initialized = .true.
!         else
!             ... the reference body, unchanged ...
!         endif
!     end subroutine
!
!     subroutine ldu_solve(solver, A, x, b)                 ! :160-176
!         call sigb_check( sigb_solver_solve(solver%dev, A%device_handle(), x, b, c_null_ptr) )
!     end subroutine
!
!     and in cg_solve_pc (cg_solvers.f90:155-194) the select type(pc) shown above gains
!         type is (sparse_ldu_solver)
!             call sigb_check( sigb_solver_solve(solver%dev, h, x, b, pc%dev) )
