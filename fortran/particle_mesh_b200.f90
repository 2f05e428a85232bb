!! particle_mesh_b200.f90 -- drop-in replacement for particle_mesh_threaded.f90 that calls the B200 library.
!! NOT compiled in this repository's CI (no Fortran compiler in the image); kept tiny on purpose.
!! Build: add to OBJS instead of particle_mesh_threaded.o, link -lcubep3m_b200 -lcudart -lnccl.
module cubep3m_b200
  use iso_c_binding
  implicit none
  type, bind(C) :: b200_config
    integer(c_int32_t) :: nodes_dim, tiles_node_dim, nf_tile, nf_buf, nf_cutoff, mesh_scale, pp_range
    integer(c_int32_t) :: max_np, max_buf, max_llf
    real(c_float)      :: density_buffer, rsoft, pp_bias, dt_pp_scale, G, eps
    integer(c_int32_t) :: ngp, ppint, pp_ext, coarse_ngp, pid, lrckcorr, move_grid_back
    integer(c_int32_t) :: ngp_fmesh_force, pp_force_flag, pp_ext_force_flag, coarse_vel_update
    integer(c_int32_t) :: rank, local_gpu, tile_split, tile_split_rank
    integer(c_int32_t) :: nodes_dim_xyz(3)
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
    integer(c_int) function b200_update_position(ctx, dt, dt_old, offset) bind(C, name='cubep3m_b200_update_position')
      import
      type(c_ptr), value :: ctx
      real(c_float), value :: dt, dt_old
      real(c_float), intent(in) :: offset(3)
    end function
  end interface
end module cubep3m_b200

!! same name, same COMMON includes as the routine it replaces (particle_mesh_threaded.f90:2-6)
subroutine particle_mesh
  use iso_c_binding
  use cubep3m_b200
  implicit none
  include 'mpif.h'
  include 'cubepm.fh'
  type(b200_config) :: cfg
  type(b200_step_out) :: o
  real(4) :: offset(3)
  integer(4) :: st, np_dl
  character(kind=c_char), target :: nccl_id(128)

  if (.not. c_associated(b200_ctx)) then            ! first call: one-time init from the compile-time parameters
    cfg%nodes_dim = nodes_dim; cfg%tiles_node_dim = tiles_node_dim; cfg%nf_tile = nf_tile
    cfg%nf_buf = nf_buf; cfg%nf_cutoff = nf_cutoff; cfg%mesh_scale = mesh_scale; cfg%pp_range = pp_range
    cfg%max_np = max_np; cfg%max_buf = max_buf; cfg%max_llf = max_llf
    cfg%density_buffer = density_buffer; cfg%rsoft = rsoft; cfg%pp_bias = pp_bias; cfg%dt_pp_scale = dt_pp_scale
    cfg%G = G; cfg%eps = eps
    cfg%ngp = 1; cfg%ppint = 1; cfg%pp_ext = 0; cfg%coarse_ngp = 0; cfg%pid = 0; cfg%lrckcorr = 1; cfg%move_grid_back = 0
#ifdef PP_EXT
    cfg%pp_ext = 1
#endif
#ifdef PID_FLAG
    cfg%pid = 1
#endif
    cfg%ngp_fmesh_force = merge(1, 0, ngp_fmesh_force); cfg%pp_force_flag = merge(1, 0, pp_force_flag)
    cfg%pp_ext_force_flag = merge(1, 0, pp_ext_force_flag); cfg%coarse_vel_update = merge(1, 0, coarse_vel_update)
    cfg%rank = rank; cfg%local_gpu = mod(rank, 8); cfg%tile_split = 1; cfg%tile_split_rank = 0
    cfg%nodes_dim_xyz = 0
    if (rank == 0) st = b200_get_unique_id(c_loc(nccl_id))
    call mpi_bcast(nccl_id, 128, mpi_character, 0, mpi_comm_world, ierr)
    ! kern_f / kern_c were filled by fine_kernel / coarse_kernel (cubepm.f90:42-45): hand them over as they are
    st = b200_init(cfg, c_null_ptr, c_null_ptr, c_loc(kern_f), c_loc(kern_c), c_loc(nccl_id), nodes, b200_ctx)
    if (st /= 0) call mpi_abort(mpi_comm_world, st, ierr)
    st = b200_upload(b200_ctx, c_loc(xv), c_null_ptr, np_local)
    if (st /= 0) call mpi_abort(mpi_comm_world, st, ierr)
  endif

  ! the shake offset is still drawn here, exactly as update_position.f90:56-63 does
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

#ifdef B200_STRICT
  st = b200_upload(b200_ctx, c_loc(xv), c_null_ptr, np_local)
#endif
  st = b200_particle_mesh(b200_ctx, dt, dt_old, a_mid, mass_p, offset, o)
  if (st /= 0) then
    write(*,*) 'rank:', rank, 'cubep3m_b200 status', st
    call mpi_abort(mpi_comm_world, st, ierr)
  endif
  np_local = o%np_local
  dt_f_acc = o%dt_f_acc; dt_pp_acc = o%dt_pp_acc; dt_pp_ext_acc = o%dt_pp_ext_acc; dt_c_acc = o%dt_c_acc
  if (rank == 0) then
    write(*,*) 'maximum timestep from fine force=', dt_f_acc
    write(*,*) 'maximum timestep from pp force=', dt_pp_acc
    write(*,*) 'sum of rho_f=', o%sum_rho_f
    write(*,*) 'sum of rho_c=', o%sum_rho_c
    write(*,*) 'maximum dt from coarse grid=', dt_c_acc
    write(*,*) 'total number of particles =', o%np_total
  endif
#ifdef B200_STRICT
  st = b200_download(b200_ctx, c_loc(xv), c_null_ptr, np_dl)
#else
  ! resident mode: xv is fetched only when the driver needs it (cubepm.f90:171-233)
  if (checkpoint_step .or. projection_step .or. halofind_step) st = b200_download(b200_ctx, c_loc(xv), c_null_ptr, np_dl)
#endif
end subroutine particle_mesh
