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
