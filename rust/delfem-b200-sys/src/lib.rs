//! Raw bindings of `include/delfem_b200.h` (mechanically derived from the header; keep in sync).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type dfb_status = i32;
pub const DFB_OK: dfb_status = 0;
pub const DFB_ERR_BAD_ARG: dfb_status = 1;
pub const DFB_ERR_PATTERN_MISS: dfb_status = 2;
pub const DFB_ERR_SINGULAR_DIAG: dfb_status = 3;
pub const DFB_ERR_CUDA: dfb_status = 4;
pub const DFB_ERR_NCCL: dfb_status = 5;
pub const DFB_ERR_INDEX_OVERFLOW: dfb_status = 6;
pub const DFB_ERR_NOT_POSITIVE: dfb_status = 7;
pub const DFB_ERR_UNSUPPORTED: dfb_status = 8;

macro_rules! opaque { ($($n:ident),*) => { $( #[repr(C)] pub struct $n { _p: [u8; 0] } )* } }
opaque!(dfb_ctx, dfb_vec, dfb_mesh, dfb_pattern, dfb_bsr, dfb_plan, dfb_precond, dfb_halo, dfb_bcflag);

pub const DFB_ELEM_LAPLACE_TRI2: i32 = 0;
pub const DFB_ELEM_COT_LAPLACE_TRI3: i32 = 1;
pub const DFB_ELEM_SOLID_LINEAR_TRI2: i32 = 2;
pub const DFB_ELEM_LAPLACE_TET: i32 = 3;
pub const DFB_ELEM_SOLID_LINEAR_TET: i32 = 4;
pub const DFB_ELEM_NEOHOOKEAN_TET: i32 = 5;
pub const DFB_PRECOND_NONE: i32 = 0;
pub const DFB_PRECOND_JACOBI: i32 = 1;
pub const DFB_PRECOND_BLOCK_JACOBI: i32 = 2;

extern "C" {
    pub fn dfb_last_error_message() -> *const c_char;
    pub fn dfb_ctx_create(device: i32, out: *mut *mut dfb_ctx) -> dfb_status;
    pub fn dfb_ctx_destroy(ctx: *mut dfb_ctx) -> dfb_status;
    pub fn dfb_ctx_sync(ctx: *mut dfb_ctx) -> dfb_status;
    pub fn dfb_ctx_set_option(ctx: *mut dfb_ctx, key: *const c_char, value: i64) -> dfb_status;
    pub fn dfb_host_register(ctx: *mut dfb_ctx, host: *mut c_void, bytes: u64) -> dfb_status;
    pub fn dfb_host_unregister(ctx: *mut dfb_ctx, host: *mut c_void) -> dfb_status;

    pub fn dfb_vec_create(ctx: *mut dfb_ctx, n_blk: u64, n_ghost: u64, blk_dim: i32, out: *mut *mut dfb_vec) -> dfb_status;
    pub fn dfb_vec_destroy(v: *mut dfb_vec) -> dfb_status;
    pub fn dfb_vec_upload(v: *mut dfb_vec, host: *const f64) -> dfb_status;
    pub fn dfb_vec_download(v: *mut dfb_vec, host: *mut f64) -> dfb_status;
    pub fn dfb_vec_set_zero(v: *mut dfb_vec) -> dfb_status;
    pub fn dfb_vec_copy(dst: *mut dfb_vec, src: *const dfb_vec) -> dfb_status;
    pub fn dfb_vec_dot(a: *const dfb_vec, b: *const dfb_vec, out: *mut f64) -> dfb_status;
    pub fn dfb_vec_add_scaled(u: *mut dfb_vec, alpha: f64, p: *const dfb_vec) -> dfb_status;
    pub fn dfb_vec_scale_and_add(p: *mut dfb_vec, beta: f64, r: *const dfb_vec) -> dfb_status;
    pub fn dfb_vec_set_fixed_dof_zero(rhs: *mut dfb_vec, blk2isfix: *const i32) -> dfb_status;

    pub fn dfb_mesh_create(ctx: *mut dfb_ctx, elem2vtx: *const usize, num_elem: u64, num_node: i32, vtx2xyz: *const f64,
                           num_vtx: u64, ndim: i32, out: *mut *mut dfb_mesh) -> dfb_status;
    pub fn dfb_mesh_update_xyz(m: *mut dfb_mesh, vtx2xyz: *const f64) -> dfb_status;
    pub fn dfb_mesh_destroy(m: *mut dfb_mesh) -> dfb_status;

    pub fn dfb_pattern_create(ctx: *mut dfb_ctx, row2idx: *const usize, idx2col: *const usize, num_blk: u64, convention: i32,
                              out: *mut *mut dfb_pattern) -> dfb_status;
    pub fn dfb_pattern_from_uniform_mesh(ctx: *mut dfb_ctx, mesh: *const dfb_mesh, is_self: i32, num_rows: u64,
                                         out: *mut *mut dfb_pattern) -> dfb_status;
    pub fn dfb_pattern_download(p: *const dfb_pattern, row2idx: *mut usize, idx2col: *mut usize) -> dfb_status;
    pub fn dfb_pattern_destroy(p: *mut dfb_pattern) -> dfb_status;

    pub fn dfb_bsr_create(ctx: *mut dfb_ctx, pattern: *mut dfb_pattern, blk_dim: i32, out: *mut *mut dfb_bsr) -> dfb_status;
    pub fn dfb_bsr_destroy(m: *mut dfb_bsr) -> dfb_status;
    pub fn dfb_bsr_set_zero(m: *mut dfb_bsr) -> dfb_status;
    pub fn dfb_bsr_upload(m: *mut dfb_bsr, row2val: *const f64, idx2val: *const f64) -> dfb_status;
    pub fn dfb_bsr_download(m: *const dfb_bsr, row2val: *mut f64, idx2val: *mut f64) -> dfb_status;
    pub fn dfb_bsr_set_fixed_dof(m: *mut dfb_bsr, val_dia: f64, blk2isfix: *const i32) -> dfb_status;
    pub fn dfb_bsr_mult_vec(m: *const dfb_bsr, y: *mut dfb_vec, beta: f64, alpha: f64, x: *mut dfb_vec) -> dfb_status;
    pub fn dfb_bsr_merge_one(m: *mut dfb_bsr, emat: *const f64, node2vtx: *const usize, n_node: i32, flavour: i32) -> dfb_status;

    pub fn dfb_plan_create(ctx: *mut dfb_ctx, pattern: *mut dfb_pattern, mesh: *const dfb_mesh, flavour: i32,
                           out: *mut *mut dfb_plan) -> dfb_status;
    pub fn dfb_plan_destroy(p: *mut dfb_plan) -> dfb_status;
    pub fn dfb_plan_download_elem2nz(p: *const dfb_plan, elem2nz: *mut u32) -> dfb_status;
    pub fn dfb_assemble(plan: *mut dfb_plan, kind: i32, mesh: *const dfb_mesh, p0: f64, p1: f64, m: *mut dfb_bsr) -> dfb_status;
    pub fn dfb_emat_one(ctx: *mut dfb_ctx, kind: i32, node2xyz: *const f64, p0: f64, p1: f64, emat: *mut f64) -> dfb_status;

    pub fn dfb_precond_create(ctx: *mut dfb_ctx, kind: i32, m: *const dfb_bsr, out: *mut *mut dfb_precond) -> dfb_status;
    pub fn dfb_precond_update(pc: *mut dfb_precond, m: *const dfb_bsr) -> dfb_status;
    pub fn dfb_precond_apply(pc: *const dfb_precond, v: *mut dfb_vec) -> dfb_status;
    pub fn dfb_precond_destroy(pc: *mut dfb_precond) -> dfb_status;
    pub fn dfb_cg(m: *const dfb_bsr, r: *mut dfb_vec, u: *mut dfb_vec, ap: *mut dfb_vec, p: *mut dfb_vec, tol: f64, max_it: u64,
                  eps_sq: f64, check_pap: i32, hist_out: *mut f64, hist_cap: u64, n_hist: *mut u64) -> dfb_status;
    pub fn dfb_pcg(m: *const dfb_bsr, pc: *const dfb_precond, r: *mut dfb_vec, x: *mut dfb_vec, pr: *mut dfb_vec, p: *mut dfb_vec,
                   tol: f64, max_it: u64, eps_sq: f64, hist_out: *mut f64, hist_cap: u64, n_hist: *mut u64) -> dfb_status;
}
