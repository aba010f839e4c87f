//! Safe wrapper keeping the reference signatures (del-fem-cpu/src/sparse_square.rs, del-fem-ls/src/linearsystem.rs).
//! Non-zero `dfb_status` maps to `panic!` exactly where the reference asserts / unwraps.
use delfem_b200_sys as sys;
use std::ffi::CStr;

fn check(st: sys::dfb_status) {
    if st != sys::DFB_OK {
        let msg = unsafe { CStr::from_ptr(sys::dfb_last_error_message()) }.to_string_lossy().into_owned();
        panic!("delfem_b200: status {st}: {msg}");
    }
}

pub struct Ctx(*mut sys::dfb_ctx);
impl Ctx {
    pub fn new(device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::dfb_ctx_create(device, &mut h) });
        Ctx(h)
    }
}
impl Drop for Ctx {
    fn drop(&mut self) { unsafe { sys::dfb_ctx_destroy(self.0) }; }
}

pub struct DevVec<const N: usize> { h: *mut sys::dfb_vec, len: usize }
impl<const N: usize> DevVec<N> {
    pub fn from_slice(ctx: &Ctx, v: &[[f64; N]]) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::dfb_vec_create(ctx.0, v.len() as u64, 0, N as i32, &mut h) });
        check(unsafe { sys::dfb_vec_upload(h, v.as_ptr() as *const f64) });
        DevVec { h, len: v.len() }
    }
    pub fn to_vec(&self) -> Vec<[f64; N]> {
        let mut out = vec![[0f64; N]; self.len];
        check(unsafe { sys::dfb_vec_download(self.h, out.as_mut_ptr() as *mut f64) });
        out
    }
}
impl<const N: usize> Drop for DevVec<N> {
    fn drop(&mut self) { unsafe { sys::dfb_vec_destroy(self.h) }; }
}

/// `sparse_square::Matrix<[f64; NN]>` (del-fem-cpu/src/sparse_square.rs:75-214) resident on the device.
pub struct Matrix<const N: usize, const NN: usize> {
    pub num_blk: usize,
    pat: *mut sys::dfb_pattern,
    h: *mut sys::dfb_bsr,
}

impl<const N: usize, const NN: usize> Matrix<N, NN> {
    /// `Matrix::from_vtx2vtx(vtx2idx, idx2vtx)` (:117)
    pub fn from_vtx2vtx(ctx: &Ctx, vtx2idx: &[usize], idx2vtx: &[usize]) -> Self {
        assert_eq!(NN, N * N);
        let num_blk = vtx2idx.len() - 1;
        assert_eq!(vtx2idx[num_blk], idx2vtx.len());
        let (mut pat, mut h) = (std::ptr::null_mut(), std::ptr::null_mut());
        check(unsafe { sys::dfb_pattern_create(ctx.0, vtx2idx.as_ptr(), idx2vtx.as_ptr(), num_blk as u64, 0, &mut pat) });
        check(unsafe { sys::dfb_bsr_create(ctx.0, pat, N as i32, &mut h) });
        Matrix { num_blk, pat, h }
    }
    /// `set_zero` (:174)
    pub fn set_zero(&mut self) { check(unsafe { sys::dfb_bsr_set_zero(self.h) }); }
    /// `merge_for_array_blk` (:180-214); `col2idx` kept for signature parity (unused: the device merge needs no scratch)
    pub fn merge_for_array_blk<const NNODE: usize>(&mut self, emat: &[[[f64; NN]; NNODE]; NNODE], node2vtx: &[usize; NNODE],
                                                   _col2idx: &mut Vec<usize>) {
        check(unsafe { sys::dfb_bsr_merge_one(self.h, emat.as_ptr() as *const f64, node2vtx.as_ptr(), NNODE as i32, 0) });
    }
    /// `set_fixed_dof` (:129)
    pub fn set_fixed_dof(&mut self, val_dia: f64, blk2isfix: &[[i32; N]]) {
        assert_eq!(blk2isfix.len(), self.num_blk);
        check(unsafe { sys::dfb_bsr_set_fixed_dof(self.h, val_dia, blk2isfix.as_ptr() as *const i32) });
    }
    /// `mult_vec` (:143-171): y <- alpha*A*x + beta*y
    pub fn mult_vec(&self, y_vec: &mut DevVec<N>, beta: f64, alpha: f64, x_vec: &mut DevVec<N>) {
        assert_eq!(y_vec.len, self.num_blk);
        check(unsafe { sys::dfb_bsr_mult_vec(self.h, y_vec.h, beta, alpha, x_vec.h) });
    }
}
impl<const N: usize, const NN: usize> Drop for Matrix<N, NN> {
    fn drop(&mut self) { unsafe { sys::dfb_bsr_destroy(self.h); sys::dfb_pattern_destroy(self.pat); } }
}

/// `conjugate_gradient(r, u, ap, p, tol, max_it, mat) -> Vec<T>` (del-fem-cpu/src/sparse_square.rs:218-261)
pub fn conjugate_gradient<const N: usize, const NN: usize>(r_vec: &mut DevVec<N>, u_vec: &mut DevVec<N>, ap_vec: &mut DevVec<N>,
                                                           p_vec: &mut DevVec<N>, conv_ratio_tol: f64, max_iteration: usize,
                                                           mat: &Matrix<N, NN>) -> Vec<f64> {
    let mut hist = vec![0f64; max_iteration + 2];
    let mut n: u64 = 0;
    check(unsafe { sys::dfb_cg(mat.h, r_vec.h, u_vec.h, ap_vec.h, p_vec.h, conv_ratio_tol, max_iteration as u64, f64::EPSILON, 0,
                               hist.as_mut_ptr(), hist.len() as u64, &mut n) });
    hist.truncate(n as usize);
    hist
}

/// `set_fix_dof_to_rhs_vector` (del-fem-cpu/src/sparse_square.rs:57-70)
pub fn set_fix_dof_to_rhs_vector<const N: usize>(blk2rhs: &mut DevVec<N>, blk2isfix: &[[i32; N]]) {
    assert_eq!(blk2isfix.len(), blk2rhs.len);
    check(unsafe { sys::dfb_vec_set_fixed_dof_zero(blk2rhs.h, blk2isfix.as_ptr() as *const i32) });
}

impl<const N: usize> DevVec<N> {
    /// zero vector of `num_blk` blocks (the reference's `vec![[0.; N]; num_blk]`)
    pub fn zeros(ctx: &Ctx, num_blk: usize) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::dfb_vec_create(ctx.0, num_blk as u64, 0, N as i32, &mut h) });
        check(unsafe { sys::dfb_vec_set_zero(h) });
        DevVec { h, len: num_blk }
    }
}

/// Element kinds of the batched assembly (`dfb_assemble`): each replaces one per-element function of the reference.
#[derive(Clone, Copy)]
pub enum ElemKind {
    LaplaceTri2 = sys::DFB_ELEM_LAPLACE_TRI2 as isize,          // laplace_tri2::ddw_            laplace_tri2.rs:3
    CotLaplaceTri3 = sys::DFB_ELEM_COT_LAPLACE_TRI3 as isize,   // tri3::emat_cotangent_laplacian laplace_tri3.rs:18
    SolidLinearTri2 = sys::DFB_ELEM_SOLID_LINEAR_TRI2 as isize, // solid_linear_tri2::emat_tri2   solid_linear_tri2.rs:23
    LaplaceTet = sys::DFB_ELEM_LAPLACE_TET as isize,
    SolidLinearTet = sys::DFB_ELEM_SOLID_LINEAR_TET as isize,
    MassTet = sys::DFB_ELEM_MASS_TET as isize,
}

/// A uniform mesh resident on the device: `elem2vtx` (flat, `NNODE` per element) and `vtx2xyz` (flat, `NDIM` per vertex).
pub struct Mesh { h: *mut sys::dfb_mesh }
impl Mesh {
    pub fn new(ctx: &Ctx, elem2vtx: &[usize], num_node: usize, vtx2xyz: &[f64], num_dim: usize) -> Self {
        assert_eq!(elem2vtx.len() % num_node, 0);
        assert_eq!(vtx2xyz.len() % num_dim, 0);
        let mut h = std::ptr::null_mut();
        check(unsafe {
            sys::dfb_mesh_create(ctx.0, elem2vtx.as_ptr(), (elem2vtx.len() / num_node) as u64, num_node as i32, vtx2xyz.as_ptr(),
                                 (vtx2xyz.len() / num_dim) as u64, num_dim as i32, &mut h)
        });
        Mesh { h }
    }
}
impl Drop for Mesh {
    fn drop(&mut self) { unsafe { sys::dfb_mesh_destroy(self.h) }; }
}

/// The element -> nonzero map of (pattern, mesh), inverted once: what the reference recomputes per element with `col2idx`.
pub struct Plan { h: *mut sys::dfb_plan }
impl Plan {
    pub fn new<const N: usize, const NN: usize>(ctx: &Ctx, mat: &Matrix<N, NN>, mesh: &Mesh) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::dfb_plan_create(ctx.0, mat.pat, mesh.h, 0, &mut h) });
        Plan { h }
    }
}
impl Drop for Plan {
    fn drop(&mut self) { unsafe { sys::dfb_plan_destroy(self.h) }; }
}

impl<const N: usize, const NN: usize> Matrix<N, NN> {
    /// The whole `for node2vtx in elem2vtx.chunks(NNODE) { emat = ...; merge_for_array_blk(..) }` loop
    /// (solid_linear_tri2.rs:144-161) as one deterministic gather on the device. Overwrites the values (no `set_zero` needed).
    pub fn assemble(&mut self, plan: &Plan, kind: ElemKind, mesh: &Mesh, p0: f64, p1: f64) {
        check(unsafe { sys::dfb_assemble(plan.h, kind as i32, mesh.h, p0, p1, self.h) });
    }
    /// inertia / damping on the diagonal: `row2val[i][d + N d] += scale * v[i][d]` (rod3_darboux.rs:1219-1226)
    pub fn add_to_diagonal(&mut self, scale: f64, v: &DevVec<N>) {
        check(unsafe { sys::dfb_bsr_add_to_diagonal(self.h, scale, v.h) });
    }
}

/// `sparse_ilu::Preconditioner<[f64; NN]>` (del-fem-cpu/src/sparse_ilu.rs) on the device.
pub mod sparse_ilu {
    use super::*;
    pub struct Preconditioner<const N: usize, const NN: usize> {
        pub(crate) h: *mut sys::dfb_precond,
        src: *const sys::dfb_bsr,
    }
    impl<const N: usize, const NN: usize> Preconditioner<N, NN> {
        pub fn new() -> Self { Preconditioner { h: std::ptr::null_mut(), src: std::ptr::null() } }
        fn reset(&mut self) {
            if !self.h.is_null() { unsafe { sys::dfb_precond_destroy(self.h) }; }
            self.h = std::ptr::null_mut();
        }
        /// `initialize_ilu0` (:28-56)
        pub fn initialize_ilu0(&mut self, ctx: &Ctx, a: &Matrix<N, NN>) {
            self.reset();
            check(unsafe { sys::dfb_precond_create(ctx.0, sys::DFB_PRECOND_ILU0, a.h, &mut self.h) });
        }
        /// `initialize_iluk` (:58-64)
        pub fn initialize_iluk(&mut self, ctx: &Ctx, a: &Matrix<N, NN>, lev_fill: usize) {
            self.reset();
            check(unsafe { sys::dfb_precond_create_iluk(ctx.0, a.h, lev_fill as u64, &mut self.h) });
        }
        /// point Jacobi — the fast preconditioner on the GPU (no reference counterpart; the ILU code with an empty pattern
        /// is block-Jacobi: `DFB_PRECOND_BLOCK_JACOBI`)
        pub fn initialize_jacobi(&mut self, ctx: &Ctx, a: &Matrix<N, NN>) {
            self.reset();
            check(unsafe { sys::dfb_precond_create(ctx.0, sys::DFB_PRECOND_JACOBI, a.h, &mut self.h) });
        }
    }
    impl<const N: usize, const NN: usize> Drop for Preconditioner<N, NN> {
        fn drop(&mut self) { self.reset(); }
    }
    /// `copy_value` (:87-125) — recorded here, performed together with `decompose` on the device
    pub fn copy_value<const N: usize, const NN: usize>(ilu: &mut Preconditioner<N, NN>, a: &Matrix<N, NN>) { ilu.src = a.h; }
    /// `decompose` (:127-181); a singular pivot block panics like `try_inverse().unwrap()`
    pub fn decompose<const N: usize, const NN: usize>(ilu: &mut Preconditioner<N, NN>) {
        assert!(!ilu.src.is_null(), "decompose: call copy_value first");
        check(unsafe { sys::dfb_precond_update(ilu.h, ilu.src) });
    }
    /// `solve_preconditioning_vec` (:183-219), in place
    pub fn solve_preconditioning_vec<const N: usize, const NN: usize>(vec: &mut DevVec<N>, ilu: &Preconditioner<N, NN>) {
        check(unsafe { sys::dfb_precond_apply(ilu.h, vec.h) });
    }
}

/// `preconditioned_conjugate_gradient` (del-fem-cpu/src/sparse_square.rs:264-347, any N): history of absolute residual norms
/// including k = 0, one trailing entry when `max_nitr` is exhausted.
pub fn preconditioned_conjugate_gradient<const N: usize, const NN: usize>(
    r_vec: &mut DevVec<N>, x_vec: &mut DevVec<N>, pr_vec: &mut DevVec<N>, p_vec: &mut DevVec<N>, conv_ratio_tol: f64, max_nitr: usize,
    mat: &Matrix<N, NN>, ilu: &sparse_ilu::Preconditioner<N, NN>) -> Vec<f64> {
    let mut hist = vec![0f64; max_nitr + 3];
    let mut n: u64 = 0;
    check(unsafe { sys::dfb_pcg(mat.h, ilu.h, r_vec.h, x_vec.h, pr_vec.h, p_vec.h, conv_ratio_tol, max_nitr as u64, f64::EPSILON,
                                hist.as_mut_ptr(), hist.len() as u64, &mut n) });
    hist.truncate(n as usize);
    hist
}

/// `del_fem_ls::linearsystem::Solver` (del-fem-ls/src/linearsystem.rs:8-112) for block size N over `dfb_solver_*`: same
/// lifecycle (`initialize` -> `begin_merge` -> merges -> `end_merge` -> `solve_cg` / `solve_pcg`), same defaults (tol 1e-5f32,
/// 100 iterations, ILU(0) set up by `initialize` as :53), `clone` resets `max_num_iteration` to 100 as :109 does.
/// `sparse()`, `r_vec()`, `u_vec()`, `ilu()` are the reference's pub fields as handles BORROWED from the solver.
pub struct Solver<const N: usize> { h: *mut sys::dfb_solver, pub conv: Vec<f64> }
impl<const N: usize> Solver<N> {
    /// `Solver::new` (:33)
    pub fn new(ctx: &Ctx) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::dfb_solver_new(ctx.0, N as i32, &mut h) });
        Solver { h, conv: Vec::new() }
    }
    /// `initialize(colind, rowptr)` (:48): first slice = row pointer, second = column indices (names swapped upstream)
    pub fn initialize(&mut self, colind: &[usize], rowptr: &[usize]) {
        check(unsafe { sys::dfb_solver_initialize(self.h, colind.as_ptr() as *const u64, colind.len() as u64,
                                                  rowptr.as_ptr() as *const u64, rowptr.len() as u64) });
    }
    pub fn begin_merge(&mut self) { check(unsafe { sys::dfb_solver_begin_merge(self.h) }); }   // :60
    pub fn end_merge(&mut self) { check(unsafe { sys::dfb_solver_end_merge(self.h) }); }       // :67
    pub fn solve_cg(&mut self) { check(unsafe { sys::dfb_solver_solve_cg(self.h) }); self.fetch_conv(); }    // :73
    pub fn solve_pcg(&mut self) { check(unsafe { sys::dfb_solver_solve_pcg(self.h) }); self.fetch_conv(); }  // :85
    pub fn sparse(&self) -> *mut sys::dfb_bsr { unsafe { sys::dfb_solver_sparse(self.h) } }
    pub fn ilu(&self) -> *mut sys::dfb_precond { unsafe { sys::dfb_solver_ilu(self.h) } }
    pub fn r_vec(&self) -> *mut sys::dfb_vec { unsafe { sys::dfb_solver_r_vec(self.h) } }
    pub fn u_vec(&self) -> *mut sys::dfb_vec { unsafe { sys::dfb_solver_u_vec(self.h) } }
    pub fn set_params(&mut self, conv_ratio_tol: f64, max_num_iteration: usize) {
        check(unsafe { sys::dfb_solver_set_params(self.h, conv_ratio_tol, max_num_iteration as u64) });
    }
    fn fetch_conv(&mut self) {
        let mut n = 0u64;
        check(unsafe { sys::dfb_solver_conv(self.h, std::ptr::null_mut(), 0, &mut n) });
        self.conv = vec![0.0; n as usize];
        check(unsafe { sys::dfb_solver_conv(self.h, self.conv.as_mut_ptr(), n, &mut n) });
    }
}
impl<const N: usize> Clone for Solver<N> {
    /// `Solver::clone` (:96-111)
    fn clone(&self) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::dfb_solver_clone(self.h, &mut h) });
        let mut t = Solver { h, conv: Vec::new() };
        t.fetch_conv();
        t
    }
}
impl<const N: usize> Drop for Solver<N> {
    fn drop(&mut self) { unsafe { sys::dfb_solver_destroy(self.h) }; }
}
