//! UNVERIFIED -- written against include/tfhe_b200.h, never compiled (no Rust toolchain in the build image).
//!
//! Safe wrappers named like the functions of Janmajayamall/tfhe-research that they replace
//! (`bootstrap` bootstrapping.rs:58, `and`/`or` boolean.rs:9-53, `key_switch_lwe` key_switching.rs:63,
//! `external_product`/`cmux` ggsw.rs:132-178), over flat `u32` slices in exactly the ndarray layouts of the
//! reference (`Array1<u32>[n+1]`, `Array2<u32>[k+1,N]`, `Array3<u32>[(k+1)l,k+1,N]`, ...).
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct TfheParams {
    pub glwe_dimension: u32,
    pub glwe_poly_degree: u32, // log2 N, as in lib.rs:23-34
    pub lwe_dimension: u32,
    pub padding_bits: u32,
    pub log_p: u32,
    pub log_q: u32,
    pub ks_log_base: u32,
    pub ks_levels: u32,
    pub pbs_log_base: u32,
    pub pbs_levels: u32,
    pub lwe_std_dev: f64,
    pub glwe_std_dev: f64,
}
#[repr(C)] pub struct RawCtx { _p: [u8; 0] }
#[repr(C)] pub struct RawBk { _p: [u8; 0] }

extern "C" {
    fn tfhe_params_default(test_cfg: c_int, out: *mut TfheParams) -> c_int;
    fn tfhe_ctx_create(p: *const TfheParams, device: c_int, out: *mut *mut RawCtx) -> c_int;
    fn tfhe_ctx_destroy(ctx: *mut RawCtx);
    fn tfhe_last_error(ctx: *const RawCtx) -> *const c_char;
    fn tfhe_ctx_set_pbs_path(ctx: *mut RawCtx, path: c_int) -> c_int;
    fn tfhe_ctx_set_ks_path(ctx: *mut RawCtx, path: c_int) -> c_int;
    fn tfhe_bk_upload(ctx: *mut RawCtx, bsk: *const u32, ksk: *const u32, out: *mut *mut RawBk) -> c_int;
    fn tfhe_bk_free(bk: *mut RawBk);
    fn tfhe_bootstrap_batch(ctx: *mut RawCtx, bk: *const RawBk, lwe_in: *const u32, luts: *const u32, n_luts: usize,
                            lut_idx: *const u32, batch: usize, lwe_out: *mut u32) -> c_int;
    fn tfhe_bootstrap_batch_ks_first(ctx: *mut RawCtx, bk: *const RawBk, lwe_in: *const u32, luts: *const u32, n_luts: usize,
                                     lut_idx: *const u32, batch: usize, lwe_out: *mut u32) -> c_int;
    fn tfhe_gate_batch(ctx: *mut RawCtx, bk: *const RawBk, gate: c_int, ct0: *const u32, ct1: *const u32, batch: usize,
                       out: *mut u32) -> c_int;
    fn tfhe_external_product(ctx: *mut RawCtx, bk: *const RawBk, ggsw_index: *const u32, glwe: *const u32, batch: usize,
                             out: *mut u32) -> c_int;
    fn tfhe_cmux(ctx: *mut RawCtx, bk: *const RawBk, ggsw_index: *const u32, ct0: *const u32, ct1: *const u32, batch: usize,
                 out: *mut u32) -> c_int;
    fn tfhe_key_switch(ctx: *mut RawCtx, bk: *const RawBk, lwe_in: *const u32, batch: usize, lwe_out: *mut u32) -> c_int;
    fn tfhe_sample_extract(ctx: *mut RawCtx, glwe: *const u32, batch: usize, lwe_out: *mut u32) -> c_int;
    fn tfhe_test_vector_identity(p: *const TfheParams, tv_out: *mut u32) -> c_int;
    fn tfhe_test_vector_boolean(p: *const TfheParams, gate: c_int, tv_out: *mut u32) -> c_int;
}

#[derive(Debug)]
pub struct Error { pub code: i32, pub message: String }

#[repr(i32)]
#[derive(Clone, Copy)]
pub enum Gate { And = 0, Or = 1, Xor = 2, Nand = 3, Nor = 4, Xnor = 5 }

impl Default for TfheParams {
    /// lib.rs:101-123 (the non-test `impl Default for TfheParams`)
    fn default() -> Self {
        let mut p = std::mem::MaybeUninit::<TfheParams>::uninit();
        unsafe { assert_eq!(tfhe_params_default(0, p.as_mut_ptr()), 0); p.assume_init() }
    }
}

/// One B200 + a device-resident `BootstrappingKey` (bootstrapping.rs:18-21).
pub struct B200 { ctx: *mut RawCtx, bk: *mut RawBk, pub params: TfheParams }

impl B200 {
    /// `bsk` = concatenation of `bk.lwe_sk_ggsw_enc[i].data` (Array3<u32> in standard layout), `ksk` = `bk.ksk.data`.
    pub fn new(params: TfheParams, device: i32, bsk: &[u32], ksk: &[u32]) -> Result<Self, Error> {
        let (n, k, l) = (params.lwe_dimension as usize, params.glwe_dimension as usize, params.pbs_levels as usize);
        let big_n = 1usize << params.glwe_poly_degree;
        assert_eq!(bsk.len(), n * (k + 1) * l * (k + 1) * big_n);
        assert_eq!(ksk.len(), k * big_n * params.ks_levels as usize * (n + 1));
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { tfhe_ctx_create(&params, device, &mut ctx) };
        if rc != 0 { return Err(Error { code: rc, message: "tfhe_ctx_create (no CUDA device? there is no CPU fallback)".into() }); }
        let mut me = B200 { ctx, bk: std::ptr::null_mut(), params };
        let rc = unsafe { tfhe_bk_upload(me.ctx, bsk.as_ptr(), ksk.as_ptr(), &mut me.bk) };
        me.check(rc)?;
        Ok(me)
    }
    fn check(&self, rc: c_int) -> Result<(), Error> {
        if rc == 0 { return Ok(()); }
        let msg = unsafe { CStr::from_ptr(tfhe_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message: msg })    // -4 (TFHE_E_ASSERT) == a reference assert! would have fired
    }
    fn lwe_words(&self) -> usize { self.params.lwe_dimension as usize + 1 }
    fn glwe_words(&self) -> usize { (self.params.glwe_dimension as usize + 1) << self.params.glwe_poly_degree }

    /// drop-in for `bootstrap` (bootstrapping.rs:58-120) over a batch; `test_vector_poly` is UNENCODED like :84.
    pub fn bootstrap(&self, lwe_cts: &[u32], test_vector_poly: &[u32]) -> Result<Vec<u32>, Error> {
        let batch = lwe_cts.len() / self.lwe_words();
        let mut out = vec![0u32; lwe_cts.len()];
        let rc = unsafe { tfhe_bootstrap_batch(self.ctx, self.bk, lwe_cts.as_ptr(), test_vector_poly.as_ptr(), 1,
                                               std::ptr::null(), batch, out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out)
    }
    /// key switch first, then blind rotation + sample extraction (notes/TFHE.md:365-400); dimension kN in and out.
    pub fn bootstrap_ks_first(&self, lwe_cts: &[u32], test_vector_poly: &[u32]) -> Result<Vec<u32>, Error> {
        let words = ((self.params.glwe_dimension as usize) << self.params.glwe_poly_degree) + 1;
        let batch = lwe_cts.len() / words;
        let mut out = vec![0u32; lwe_cts.len()];
        let rc = unsafe { tfhe_bootstrap_batch_ks_first(self.ctx, self.bk, lwe_cts.as_ptr(), test_vector_poly.as_ptr(), 1,
                                                        std::ptr::null(), batch, out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out)
    }
    /// drop-in for `and` (boolean.rs:9-30) / `or` (boolean.rs:32-53) and the gates defined by this build.
    pub fn gate(&self, gate: Gate, ct0: &[u32], ct1: &[u32]) -> Result<Vec<u32>, Error> {
        assert_eq!(ct0.len(), ct1.len());
        let batch = ct0.len() / self.lwe_words();
        let mut out = vec![0u32; ct0.len()];
        let rc = unsafe { tfhe_gate_batch(self.ctx, self.bk, gate as c_int, ct0.as_ptr(), ct1.as_ptr(), batch, out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out)
    }
    pub fn and(&self, ct0: &[u32], ct1: &[u32]) -> Result<Vec<u32>, Error> { self.gate(Gate::And, ct0, ct1) }
    pub fn or(&self, ct0: &[u32], ct1: &[u32]) -> Result<Vec<u32>, Error> { self.gate(Gate::Or, ct0, ct1) }
    /// drop-in for `external_product` (ggsw.rs:132-161); the GGSW is addressed by its index in the uploaded BSK.
    pub fn external_product(&self, ggsw_index: &[u32], glwe: &[u32]) -> Result<Vec<u32>, Error> {
        let mut out = vec![0u32; glwe.len()];
        let rc = unsafe { tfhe_external_product(self.ctx, self.bk, ggsw_index.as_ptr(), glwe.as_ptr(), ggsw_index.len(), out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out)
    }
    /// drop-in for `cmux` (ggsw.rs:164-178); unlike the reference `ct1` is not clobbered.
    pub fn cmux(&self, ggsw_index: &[u32], ct0: &[u32], ct1: &[u32]) -> Result<Vec<u32>, Error> {
        let mut out = vec![0u32; ct0.len()];
        let rc = unsafe { tfhe_cmux(self.ctx, self.bk, ggsw_index.as_ptr(), ct0.as_ptr(), ct1.as_ptr(), ggsw_index.len(), out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out)
    }
    /// drop-in for `key_switch_lwe` (key_switching.rs:63-103): [B][kN+1] -> [B][n+1]
    pub fn key_switch_lwe(&self, lwe_in: &[u32]) -> Result<Vec<u32>, Error> {
        let words = ((self.params.glwe_dimension as usize) << self.params.glwe_poly_degree) + 1;
        let batch = lwe_in.len() / words;
        let mut out = vec![0u32; batch * self.lwe_words()];
        let rc = unsafe { tfhe_key_switch(self.ctx, self.bk, lwe_in.as_ptr(), batch, out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out)
    }
    /// drop-in for `sample_extract` (bootstrapping.rs:122-156) with sample_index 0
    pub fn sample_extract(&self, glwe: &[u32]) -> Result<Vec<u32>, Error> {
        let batch = glwe.len() / self.glwe_words();
        let words = ((self.params.glwe_dimension as usize) << self.params.glwe_poly_degree) + 1;
        let mut out = vec![0u32; batch * words];
        let rc = unsafe { tfhe_sample_extract(self.ctx, glwe.as_ptr(), batch, out.as_mut_ptr()) };
        self.check(rc)?;
        Ok(out)
    }
    /// `construct_identity_test_vector` (test_vector.rs:23-35)
    pub fn construct_identity_test_vector(&self) -> Vec<u32> {
        let mut tv = vec![0u32; 1 << self.params.glwe_poly_degree];
        unsafe { assert_eq!(tfhe_test_vector_identity(&self.params, tv.as_mut_ptr()), 0) };
        tv
    }
    /// `construct_test_vector_boolean` (test_vector.rs:5-20) for AND / OR / XOR
    pub fn construct_test_vector_boolean(&self, gate: Gate) -> Vec<u32> {
        let mut tv = vec![0u32; 1 << self.params.glwe_poly_degree];
        unsafe { assert_eq!(tfhe_test_vector_boolean(&self.params, gate as c_int, tv.as_mut_ptr()), 0) };
        tv
    }
    /// 0 = 2-prime NTT path, 1 = exact FP64-FFT path; call before uploading a key (i.e. before `new` returns -- see
    /// tfhe_ctx_set_pbs_path; exposed here for completeness).
    pub fn set_pbs_path(&self, path: i32) -> Result<(), Error> { let rc = unsafe { tfhe_ctx_set_pbs_path(self.ctx, path) }; self.check(rc) }
    /// Arithmetic of the key-switching product: 0 = 32-bit multiply-adds, 1 = integer tensor cores (same bits).
    pub fn set_ks_path(&self, path: i32) -> Result<(), Error> { let rc = unsafe { tfhe_ctx_set_ks_path(self.ctx, path) }; self.check(rc) }
}
impl Drop for B200 {
    fn drop(&mut self) { unsafe { if !self.bk.is_null() { tfhe_bk_free(self.bk); } tfhe_ctx_destroy(self.ctx); } }
}

// ------------------------------------------------------------------ all GPUs of the box behind one handle (tfhe_mgpu_*)
#[repr(C)] pub struct RawMgpu { _p: [u8; 0] }
#[repr(C)] pub struct RawMgpuBk { _p: [u8; 0] }

extern "C" {
    fn tfhe_mgpu_create(p: *const TfheParams, n_gpus: c_int, devices: *const c_int, out: *mut *mut RawMgpu) -> c_int;
    fn tfhe_mgpu_destroy(m: *mut RawMgpu);
    fn tfhe_mgpu_last_error(m: *const RawMgpu) -> *const c_char;
    fn tfhe_mgpu_bk_upload(m: *mut RawMgpu, bsk: *const u32, ksk: *const u32, out: *mut *mut RawMgpuBk) -> c_int;
    fn tfhe_mgpu_bk_free(bk: *mut RawMgpuBk);
    fn tfhe_mgpu_bootstrap_batch(m: *mut RawMgpu, bk: *const RawMgpuBk, lwe_in: *const u32, luts: *const u32, n_luts: usize,
                                 lut_idx: *const u32, batch: usize, lwe_out: *mut u32) -> c_int;
    fn tfhe_mgpu_gates_batch(m: *mut RawMgpu, bk: *const RawMgpuBk, gates: *const u8, ct0: *const u32, ct1: *const u32,
                             batch: usize, out: *mut u32) -> c_int;
}

/// `B200` over every GPU of the box: batches are split into contiguous balanced ranges, keys replicated.
pub struct B200Multi { m: *mut RawMgpu, bk: *mut RawMgpuBk, n: usize }

impl B200Multi {
    pub fn new(params: TfheParams, n_gpus: i32, bsk: &[u32], ksk: &[u32]) -> Result<Self, Error> {
        let mut m = std::ptr::null_mut();
        let rc = unsafe { tfhe_mgpu_create(&params, n_gpus, std::ptr::null(), &mut m) };
        if rc != 0 { return Err(Error { code: rc, message: "tfhe_mgpu_create (no CUDA device? there is no CPU fallback)".into() }); }
        let mut bk = std::ptr::null_mut();
        let rc = unsafe { tfhe_mgpu_bk_upload(m, bsk.as_ptr(), ksk.as_ptr(), &mut bk) };
        if rc != 0 {
            let message = unsafe { CStr::from_ptr(tfhe_mgpu_last_error(m)) }.to_string_lossy().into_owned();
            unsafe { tfhe_mgpu_destroy(m) };
            return Err(Error { code: rc, message });
        }
        Ok(B200Multi { m, bk, n: params.lwe_dimension as usize })
    }
    fn check(&self, rc: c_int) -> Result<(), Error> {
        if rc == 0 { Ok(()) } else { Err(Error { code: rc, message: unsafe { CStr::from_ptr(tfhe_mgpu_last_error(self.m)) }.to_string_lossy().into_owned() }) }
    }
    /// `bootstrap` (bootstrapping.rs:58) over a batch of flat LWE ciphertexts, sharded over all GPUs.
    pub fn bootstrap(&self, lwe_cts: &[u32], test_vector_poly: &[u32]) -> Result<Vec<u32>, Error> {
        let batch = lwe_cts.len() / (self.n + 1);
        let mut out = vec![0u32; lwe_cts.len()];
        let rc = unsafe { tfhe_mgpu_bootstrap_batch(self.m, self.bk, lwe_cts.as_ptr(), test_vector_poly.as_ptr(), 1, std::ptr::null(), batch, out.as_mut_ptr()) };
        self.check(rc).map(|_| out)
    }
    /// one gate opcode per ciphertext pair (`and`/`or` boolean.rs:9-53 + XOR/NAND/NOR/XNOR), sharded over all GPUs.
    pub fn gates(&self, gates: &[u8], ct0: &[u32], ct1: &[u32]) -> Result<Vec<u32>, Error> {
        let mut out = vec![0u32; ct0.len()];
        let rc = unsafe { tfhe_mgpu_gates_batch(self.m, self.bk, gates.as_ptr(), ct0.as_ptr(), ct1.as_ptr(), gates.len(), out.as_mut_ptr()) };
        self.check(rc).map(|_| out)
    }
}
impl Drop for B200Multi {
    fn drop(&mut self) { unsafe { tfhe_mgpu_bk_free(self.bk); tfhe_mgpu_destroy(self.m); } }
}
