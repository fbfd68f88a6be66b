//! UNCOMPILED (no Rust toolchain in the build image) -- see rust/README.md in the nutdb_b200 repository.
//!
//! `nutdb::parser::gpu`: the reference's public entry `Parser::parse` (src/parser/mod.rs:26-29) on a B200, through
//! the C ABI of `libnutdb_gpu.so` (`nutdb-gpu-sys`).  Same signature, same `Statement<'a>` borrowing from the input,
//! same `ParseError`s with the same positions.
use std::cell::RefCell;
use std::ffi::CStr;

use nutdb_gpu_sys as sys;

use crate::parser::{ParseError, Statement};

mod hydrate;

#[derive(Debug)]
pub struct GpuError(pub String);

/// One CUDA device.  Not `Sync`: one live batch per context (the library re-uses its pinned output buffers).
pub struct Context {
    raw: *mut sys::NutdbCtx,
}

impl Context {
    pub fn new(device: i32) -> Result<Context, GpuError> {
        let raw = unsafe { sys::nutdb_gpu_ctx_create(device) };
        if raw.is_null() {
            return Err(GpuError(format!("no usable CUDA device {device} (the GPU parser has no CPU fallback)")));
        }
        Ok(Context { raw })
    }

    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sys::nutdb_gpu_last_error(self.raw)).to_string_lossy().into_owned() }
    }

    /// `Parser::parse` for every statement of `sqls`, in one pass over the GPU.
    pub fn parse_batch<'a>(&mut self, sqls: &[&'a str]) -> Result<Vec<Result<Statement<'a>, ParseError>>, GpuError> {
        // one buffer + offsets (the library wants the statements contiguous; 16 readable bytes of padding)
        let total: usize = sqls.iter().map(|s| s.len()).sum();
        let mut text = Vec::with_capacity(total + 16);
        let mut offs = Vec::with_capacity(sqls.len() + 1);
        offs.push(0u64);
        for s in sqls {
            text.extend_from_slice(s.as_bytes());
            offs.push(text.len() as u64);
        }
        text.resize(total + 16, 0);
        let mut b: sys::NutdbBatch = unsafe { core::mem::zeroed() };
        let rc = unsafe {
            sys::nutdb_gpu_parse_batch(self.raw, text.as_ptr(), offs.as_ptr(), sqls.len() as u64, sys::F_NO_TOKENS | sys::F_WIRE_STMT, &mut b)
        };
        if rc != sys::OK {
            return Err(GpuError(format!("nutdb_gpu_parse_batch failed ({rc}): {}", self.last_error())));
        }
        let out = unsafe { hydrate::batch(&b, sqls) };
        unsafe { sys::nutdb_gpu_batch_free(self.raw, &mut b) };
        Ok(out)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::nutdb_gpu_ctx_destroy(self.raw) }
    }
}

thread_local! {
    static CTX: RefCell<Option<Context>> = RefCell::new(None);
}

/// Drop-in for `Parser::parse(sql)` (src/parser/mod.rs:27): one statement, first error wins.  Panics when no B200 is
/// present -- the library has no CPU fallback; callers that want one call `Parser::parse` themselves.
pub fn parse<'a>(sql: &'a str) -> Result<Statement<'a>, ParseError> {
    parse_batch(&[sql]).pop().expect("one result per statement")
}

/// `Parser::parse` applied to every element: where the GPU pays off (a batch of one costs ~20 kernel launches).
pub fn parse_batch<'a>(sqls: &[&'a str]) -> Vec<Result<Statement<'a>, ParseError>> {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            *c = Some(Context::new(0).expect("GPU parser unavailable"));
        }
        c.as_mut().unwrap().parse_batch(sqls).expect("GPU parser failed")
    })
}

/// All GPUs of one box from one process (`nutdb_gpu_mctx_*`): the batch is cut at statement boundaries into chunks,
/// contiguous ranges of chunks go to the devices, every chunk's results are gathered into pinned host memory over
/// that device's PCIe link and re-hydrated in the callback.
pub struct MultiGpu {
    raw: *mut sys::NutdbMCtx,
}

impl MultiGpu {
    pub fn new(devices: &[i32], workers_per_device: i32) -> Result<MultiGpu, GpuError> {
        let raw = unsafe { sys::nutdb_gpu_mctx_create(devices.as_ptr(), devices.len() as i32, workers_per_device) };
        if raw.is_null() {
            return Err(GpuError("cannot create the multi-GPU dispatcher".into()));
        }
        Ok(MultiGpu { raw })
    }

    /// Results in statement order.  `text` holds the statements back to back, `offs[i]..offs[i + 1]` is statement i.
    pub fn parse_batch<'a>(&mut self, text: &'a str, offs: &[u64], chunk_bytes: u64)
                           -> Result<Vec<Result<Statement<'a>, ParseError>>, GpuError> {
        struct State<'a, 'b> {
            text: &'a str,
            offs: &'b [u64],
            out: std::sync::Mutex<Vec<Option<Result<Statement<'a>, ParseError>>>>,
        }
        unsafe extern "C" fn on_chunk(user: *mut core::ffi::c_void, c: *const sys::NutdbMChunk) {
            let st = &*(user as *const State);
            let c = &*c;
            let first = c.first_stmt as usize;
            let n = c.batch.n_stmt as usize;
            let sqls: Vec<&str> = (0..n)
                .map(|i| &st.text[st.offs[first + i] as usize..st.offs[first + i + 1] as usize])
                .collect();
            let res = hydrate::batch(&c.batch, &sqls);
            let mut out = st.out.lock().unwrap();
            for (i, r) in res.into_iter().enumerate() {
                out[first + i] = Some(r);
            }
        }
        let n = offs.len() - 1;
        let st = State { text, offs, out: std::sync::Mutex::new((0..n).map(|_| None).collect()) };
        let rc = unsafe {
            sys::nutdb_gpu_mctx_parse_stream(self.raw, text.as_ptr(), offs.as_ptr(), n as u64, chunk_bytes, sys::F_NO_TOKENS | sys::F_WIRE_STMT,
                                             Some(on_chunk), &st as *const State as *mut core::ffi::c_void)
        };
        if rc != sys::OK {
            let msg = unsafe { CStr::from_ptr(sys::nutdb_gpu_mctx_last_error(self.raw)).to_string_lossy().into_owned() };
            return Err(GpuError(format!("nutdb_gpu_mctx_parse_stream failed ({rc}): {msg}")));
        }
        Ok(st.out.into_inner().unwrap().into_iter().map(|r| r.expect("every statement belongs to one chunk")).collect())
    }
}

impl Drop for MultiGpu {
    fn drop(&mut self) {
        unsafe { sys::nutdb_gpu_mctx_destroy(self.raw) }
    }
}
