!! particle_mesh_b200.f90 -- drop-in replacement for particle_mesh_threaded.f90 (and, in resident mode, for the sub-steps the driver
!! calls on its own at checkpoint steps) that calls the B200 library through ISO_C_BINDING.
!!
!! NOT compiled in this repository (no Fortran compiler / MPI in the image, SURVEY 0.1). The call order it implements is replayed by
!! tests/replay_driver.c through the same C entry points (strict and resident mode, with a checkpoint step) and checked against the oracle.
!!
!! Build: replace particle_mesh_threaded.o by particle_mesh_b200.o in OBJS (Make_PP_THREADS:16), compile with -DB200, link
!!        -lcubep3m_b200 -lcudart -lnccl.
!!   -DB200_STRICT   : upload before / download after every particle_mesh -- bit-for-bit drop-in for any code that touches xv between steps;
!!                     update_position.o, link_list.o, particle_pass.o, delete_particles.o, move_grid_back.o stay the reference's own.
!!   (default) resident: the device copy is authoritative between steps. ALSO remove update_position.o and move_grid_back.o from OBJS:
!!                     this file then provides `update_position` and `move_grid_back` (the two routines of cubepm.f90:171-233 that modify
!!                     xv), acting on the device copy and downloading afterwards. The reference's link_list / particle_pass /
!!                     delete_particles keep running on the downloaded host xv at halofind / projection steps (their consumers need the
!!                     host ll/hoc chains); they only add and remove ghosts, so the device copy's particle set stays valid.
!!                     b200_link_list / b200_particle_pass / b200_delete_particles below are the device versions for drivers that only
!!                     need the particle sets (report_force.f90:32,92).
module cubep3m_b200
  use iso_c_binding
  implicit none
  type, bind(C) :: b200_config            ! mirrors struct cubep3m_b200_config (include/cubep3m_b200.h) field by field
    integer(c_int32_t) :: nodes_dim, tiles_node_dim, nf_tile, nf_buf, nf_cutoff, mesh_scale, pp_range
    integer(c_int32_t) :: max_np, max_buf, max_llf
    real(c_float)      :: density_buffer, rsoft, pp_bias, dt_pp_scale, G, eps
    integer(c_int32_t) :: ngp, ppint, pp_ext, coarse_ngp, pid, lrckcorr, move_grid_back
    integer(c_int32_t) :: ngp_fmesh_force, pp_force_flag, pp_ext_force_flag, coarse_vel_update
    integer(c_int32_t) :: rank, local_gpu
    integer(c_int32_t) :: nodes_dim_xyz(3)
  end type
  type, bind(C) :: b200_peak               ! struct cubep3m_b200_peak: ipeak(1:3), 0-based tile, den_peak, peak_pos + offset
    integer(c_int32_t) :: i, j, k, tile
    real(c_float) :: den, x, y, z
  end type
  type, bind(C) :: b200_step_out
    integer(c_int32_t) :: np_local, np_with_ghosts, np_deleted_ll, np_buf_max
    real(c_float)      :: dt_f_acc, dt_pp_acc, dt_pp_ext_acc, dt_c_acc
    real(c_float)      :: f_force_max, pp_force_max, pp_ext_force_max, c_force_max
    real(c_double)     :: sum_rho_f, sum_rho_c
    integer(c_int64_t) :: np_total
    real(c_float)      :: stage_ms(16)
  end type
  type(c_ptr), save :: b200_ctx = c_null_ptr
  interface
    integer(c_int) function b200_init(cfg, fine_table, coarse_table, kern_f, kern_c, nccl_id, world, ctx) bind(C, name='cubep3m_b200_init')
      import
      type(b200_config), intent(in) :: cfg
      type(c_ptr), value :: fine_table, coarse_table, kern_f, kern_c, nccl_id
      integer(c_int), value :: world
      type(c_ptr), intent(out) :: ctx
    end function
    integer(c_int) function b200_finalize(ctx) bind(C, name='cubep3m_b200_finalize')
      import
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function b200_get_unique_id(id) bind(C, name='cubep3m_b200_get_unique_id')
      import
      type(c_ptr), value :: id
    end function
    integer(c_int) function b200_upload(ctx, xv, pid, np) bind(C, name='cubep3m_b200_upload_particles')
      import
      type(c_ptr), value :: ctx, xv, pid
      integer(c_int32_t), value :: np
    end function
    integer(c_int) function b200_download(ctx, xv, pid, np) bind(C, name='cubep3m_b200_download_particles')
      import
      type(c_ptr), value :: ctx, xv, pid
      integer(c_int32_t), intent(out) :: np
    end function
    integer(c_int) function b200_particle_mesh(ctx, dt, dt_old, a_mid, mass_p, offset, out) bind(C, name='cubep3m_b200_particle_mesh')
      import
      type(c_ptr), value :: ctx
      real(c_float), value :: dt, dt_old, a_mid, mass_p
      real(c_float), intent(in) :: offset(3)
      type(b200_step_out), intent(out) :: out
    end function
    integer(c_int) function b200_c_update_position(ctx, dt, dt_old, offset) bind(C, name='cubep3m_b200_update_position')
      import
      type(c_ptr), value :: ctx
      real(c_float), value :: dt, dt_old
      real(c_float), intent(in) :: offset(3)
    end function
    integer(c_int) function b200_c_move_grid_back(ctx, shake) bind(C, name='cubep3m_b200_move_grid_back')
      import
      type(c_ptr), value :: ctx
      real(c_float), intent(in) :: shake(3)
    end function
    integer(c_int) function b200_c_link_list(ctx, np_deleted) bind(C, name='cubep3m_b200_link_list')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), intent(out) :: np_deleted
    end function
    integer(c_int) function b200_c_particle_pass(ctx, np_with_ghosts) bind(C, name='cubep3m_b200_particle_pass')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), intent(out) :: np_with_ghosts
    end function
    integer(c_int) function b200_c_delete_particles(ctx, np) bind(C, name='cubep3m_b200_delete_particles')
      import
      type(c_ptr), value :: ctx
      integer(c_int32_t), intent(out) :: np
    end function
    function b200_strerror(st) bind(C, name='cubep3m_b200_strerror')
      import
      integer(c_int), value :: st
      type(c_ptr) :: b200_strerror
    end function
    integer(c_int) function b200_cic_power(ctx, shake, box, ngp_binning, k, d2, sig, nshells) bind(C, name='cubep3m_b200_cic_power')
      import
      type(c_ptr), value :: ctx
      real(c_float), intent(in) :: shake(3)
      real(c_double), value :: box
      integer(c_int32_t), value :: ngp_binning, nshells
      real(c_double), intent(out) :: k(*), d2(*), sig(*)
    end function
    !! find_halos' density + maxima pass on the device (halofind.f90:564-672)
    integer(c_int) function b200_c_halofind_peaks(ctx, mass_p, den_cut, para, ngph, peaks, max_peaks, n_peaks, cft) bind(C, name='cubep3m_b200_halofind_peaks')
      import
      type(c_ptr), value :: ctx
      real(c_float), value :: mass_p, den_cut
      integer(c_int32_t), value :: para, ngph, max_peaks
      type(b200_peak), intent(out) :: peaks(*)
      integer(c_int32_t), intent(out) :: n_peaks
      real(c_double), intent(out) :: cft(2)
    end function
  end interface
contains

  !! the reference's abort convention: print the reason, mpi_abort (particle_pass.f90:96-99,136-139, particle_mesh_threaded.f90:280-283)
  subroutine b200_check(st, where)
    include 'mpif.h'
    integer(c_int), intent(in) :: st
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: ierr_l, n
    if (st == 0) return
    call c_f_pointer(b200_strerror(st), msg, [64])
    n = 1
    do while (n < 64 .and. msg(n) /= c_null_char)
      n = n + 1
    enddo
    write(*,*) 'cubep3m_b200: ', where, ': ', msg(1:n-1), ' (status', st, ')'
    call mpi_abort(mpi_comm_world, st, ierr_l)
  end subroutine

  !! THE one place where the random mesh shake is drawn: update_position.f90:56-63 (rank 0 draws, everybody gets offset and shake_offset)
  subroutine b200_draw_offset(offset)
    include 'mpif.h'
    include 'cubepm.fh'
    real(4), intent(out) :: offset(3)
    offset = 0.0
#ifdef DISP_MESH
    if (rank == 0) then
      call random_number(offset)
      offset = (offset - 0.5) * mesh_scale * 4.0 - shake_offset
      shake_offset = shake_offset + offset
      print *, 'current shake offset:', shake_offset
    endif
    call mpi_bcast(offset, 3, mpi_real, 0, mpi_comm_world, ierr)
    call mpi_bcast(shake_offset, 3, mpi_real, 0, mpi_comm_world, ierr)
#endif
  end subroutine

  !! one-time init from the compile-time parameters and the driver's own kern_f / kern_c (cubepm.f90:42-45)
  subroutine b200_setup
    include 'mpif.h'
    include 'cubepm.fh'
    type(b200_config) :: cfg
    integer(4) :: st
    character(kind=c_char), target :: nccl_id(128)
    type(c_ptr) :: pid_ptr
    cfg%nodes_dim = nodes_dim; cfg%tiles_node_dim = tiles_node_dim; cfg%nf_tile = nf_tile
    cfg%nf_buf = nf_buf; cfg%nf_cutoff = nf_cutoff; cfg%mesh_scale = mesh_scale; cfg%pp_range = pp_range
    cfg%max_np = max_np; cfg%max_buf = max_buf; cfg%max_llf = max_llf
    cfg%density_buffer = density_buffer; cfg%rsoft = rsoft; cfg%pp_bias = pp_bias; cfg%dt_pp_scale = dt_pp_scale
    cfg%G = G; cfg%eps = eps
    cfg%ngp = 0; cfg%ppint = 0; cfg%pp_ext = 0; cfg%coarse_ngp = 0; cfg%pid = 0; cfg%lrckcorr = 0; cfg%move_grid_back = 0
#ifdef NGP
    cfg%ngp = 1
#endif
#ifdef PPINT
    cfg%ppint = 1
#endif
#ifdef PP_EXT
    cfg%pp_ext = 1
#endif
#ifdef COARSE_NGP
    cfg%coarse_ngp = 1
#endif
#ifdef PID_FLAG
    cfg%pid = 1
#endif
#ifdef LRCKCORR
    cfg%lrckcorr = 1
#endif
    cfg%ngp_fmesh_force = merge(1, 0, ngp_fmesh_force); cfg%pp_force_flag = merge(1, 0, pp_force_flag)
    cfg%pp_ext_force_flag = merge(1, 0, pp_ext_force_flag); cfg%coarse_vel_update = merge(1, 0, coarse_vel_update)
    cfg%rank = rank; cfg%local_gpu = mod(rank, 8)
    cfg%nodes_dim_xyz = 0                              ! the reference's cubic nodes_dim^3 grid
    if (rank == 0) st = b200_get_unique_id(c_loc(nccl_id))
    call mpi_bcast(nccl_id, 128, mpi_character, 0, mpi_comm_world, ierr)
    ! kern_f(3,nf_tile/2+1,nf_tile,nf_tile) and this rank's slab kern_c(3,nc_dim/2+1,nc_dim,nc_slab) were filled by fine_kernel /
    ! coarse_kernel: hand them over as they are (the library all-gathers the kern_c slabs, no ascii table needed)
    st = b200_init(cfg, c_null_ptr, c_null_ptr, c_loc(kern_f), c_loc(kern_c), c_loc(nccl_id), nodes, b200_ctx)
    call b200_check(st, 'init')
    pid_ptr = c_null_ptr
#ifdef PID_FLAG
    pid_ptr = c_loc(PID)
#endif
    st = b200_upload(b200_ctx, c_loc(xv), pid_ptr, np_local)
    call b200_check(st, 'upload_particles')
  end subroutine

  subroutine b200_fetch_particles
    include 'cubepm.fh'
    integer(4) :: st, np_dl
    type(c_ptr) :: pid_ptr
    pid_ptr = c_null_ptr
#ifdef PID_FLAG
    pid_ptr = c_loc(PID)
#endif
    st = b200_download(b200_ctx, c_loc(xv), pid_ptr, np_dl)
    call b200_check(st, 'download_particles')
    np_local = np_dl
  end subroutine

  !! device versions of the ghost bookkeeping for drivers that only need the particle sets (report_force.f90:32,92)
  subroutine b200_link_list
    integer(4) :: st, ndel
    st = b200_c_link_list(b200_ctx, ndel)
    call b200_check(st, 'link_list')
  end subroutine
  subroutine b200_particle_pass
    integer(4) :: st, npg
    st = b200_c_particle_pass(b200_ctx, npg)
    call b200_check(st, 'particle_pass')
  end subroutine
  subroutine b200_delete_particles
    include 'cubepm.fh'
    integer(4) :: st, np_dl
    st = b200_c_delete_particles(b200_ctx, np_dl)
    call b200_check(st, 'delete_particles')
    np_local = np_dl
  end subroutine
  !! Replacement for the first half of find_halos (halofind.f90:564-679): fills ipeak / den_peak / peak_pos of one tile, sorted by density as
  !! indexedsort leaves them, from the device pass over all tiles; the spherical-overdensity growth (:683-745) then runs unchanged on the host
  !! against a rho_f the driver deposits itself, or is skipped when only the peak catalogue is wanted. Called once per halofind step after
  !! link_list / particle_pass (cubepm.f90:193-198); tile_first(t) .. tile_first(t+1)-1 index the peaks of tile t (0-based).
  subroutine b200_halofind_peaks(peaks, n_peaks, cft)
    include 'cubepm.fh'
    type(b200_peak), intent(out) :: peaks(max_maxima)
    integer(4), intent(out) :: n_peaks
    real(8), intent(out) :: cft(2)
    integer(4) :: st, ngph, para
    ngph = 0
#ifdef NGPH
    ngph = 1
#endif
    para = 0
    if (para_inter_hc) para = 1
    st = b200_c_halofind_peaks(b200_ctx, mass_p, den_peak_cutoff, para, ngph, peaks, max_maxima, n_peaks, cft)
    call b200_check(st, 'halofind_peaks')   ! ECAPACITY = 'too many halos' (halofind.f90:626-629)
  end subroutine
end module cubep3m_b200

!! same name, same COMMON includes as the routine it replaces (particle_mesh_threaded.f90:2-6)
subroutine particle_mesh
  use iso_c_binding
  use cubep3m_b200
  implicit none
  include 'mpif.h'
  include 'cubepm.fh'
  type(b200_step_out) :: o
  real(4) :: offset(3)
  integer(4) :: st
  type(c_ptr) :: pid_ptr

  if (.not. c_associated(b200_ctx)) call b200_setup
  call b200_draw_offset(offset)                        ! particle_mesh_threaded.f90:56 -> update_position.f90:56-63
  pid_ptr = c_null_ptr
#ifdef PID_FLAG
  pid_ptr = c_loc(PID)
#endif
#ifdef B200_STRICT
  st = b200_upload(b200_ctx, c_loc(xv), pid_ptr, np_local)
  call b200_check(st, 'upload_particles')
#endif
  st = b200_particle_mesh(b200_ctx, dt, dt_old, a_mid, mass_p, offset, o)
  call b200_check(st, 'particle_mesh')
  np_local = o%np_local
  dt_f_acc = o%dt_f_acc; dt_pp_acc = o%dt_pp_acc; dt_c_acc = o%dt_c_acc
#ifdef PP_EXT
  dt_pp_ext_acc = o%dt_pp_ext_acc
#endif
  if (rank == 0) then
    write(*,*) 'maximum timestep from fine force=', dt_f_acc
    write(*,*) 'maximum timestep from pp force=', dt_pp_acc
#ifdef PP_EXT
    write(*,*) 'maximum timestep from pp ext force=', dt_pp_ext_acc
#endif
    write(*,*) 'sum of rho_f=', o%sum_rho_f
    write(*,*) 'sum of rho_c=', o%sum_rho_c
    write(*,*) 'maximum dt from coarse grid=', dt_c_acc
    write(*,*) 'total number of particles =', o%np_total
  endif
#ifdef MOVE_GRID_BACK
#error "-DMOVE_GRID_BACK is not supported by the B200 path: particle_mesh_threaded.f90:714-716 shifts the particles BEFORE delete_particles, which keeps un-kicked ghosts (DESIGN.md, out of scope); no maintained makefile sets it"
#endif
#ifdef B200_STRICT
  call b200_fetch_particles
#endif
end subroutine particle_mesh

#ifndef B200_STRICT
!! Resident mode: the two driver-called routines that MODIFY xv (cubepm.f90:176,179) act on the device copy and bring it to the host,
!! because their only callers are the checkpoint / projection / halofind blocks, which read xv next.
subroutine update_position
  use iso_c_binding
  use cubep3m_b200
  implicit none
  include 'mpif.h'
  include 'cubepm.fh'
  real(4) :: offset(3)
  integer(4) :: st
  if (.not. c_associated(b200_ctx)) call b200_setup
  call b200_draw_offset(offset)                        ! update_position.f90:56-63 — the reference draws a new offset here too
  st = b200_c_update_position(b200_ctx, dt, dt_old, offset)   ! update_position.f90:71 (dt_old = 0 at cubepm.f90:175)
  call b200_check(st, 'update_position')
  call b200_fetch_particles
end subroutine update_position

subroutine move_grid_back
  use iso_c_binding
  use cubep3m_b200
  implicit none
  include 'mpif.h'
  include 'cubepm.fh'
  integer(4) :: st, i
#ifdef DISP_MESH
  call mpi_bcast(shake_offset, 3, mpi_real, 0, mpi_comm_world, ierr)      ! move_grid_back.f90:17
  st = b200_c_move_grid_back(b200_ctx, shake_offset)
  call b200_check(st, 'move_grid_back')
  do i = 1, np_local                                   ! the host copy (valid only right after a download) takes the same unfused subtraction
    xv(1:3,i) = xv(1:3,i) - shake_offset(:)            ! move_grid_back.f90:20-23
  enddo
  shake_offset = 0.0
#endif
end subroutine move_grid_back
#endif
