//! UNCOMPILED (no Rust toolchain in the build image) -- see ../../README.md
//!
//! Mirrors `include/nutdb_gpu.h` declaration by declaration.  Constants (token types, node kinds, error codes, the
//! wire-node bit layout) keep the header's names without the `NUTDB_` prefix.
#![allow(non_camel_case_types, non_upper_case_globals)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct NutdbNode { pub kind: u8, pub sub: u8, pub aux: u16, pub parent: u32, pub a: u32, pub b: u32 }
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct NutdbNodeExt { pub index: u32, pub hdr: u32, pub a: u32, pub b: u32 }
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct NutdbStmt { pub status: u32, pub tok_begin: u32, pub tok_count: u32, pub node_begin: u32, pub node_count: u32, pub tok_used: u32 }
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct NutdbError { pub stmt: u32, pub cls: u16, pub code: u16, pub line: u32, pub col: u32, pub pos: u32, pub a: u32, pub b: u32, pub c: u32 }

#[repr(C)]
pub struct NutdbBatch {
    pub n_stmt: u64, pub n_tok: u64, pub n_node: u64, pub n_err: u64,
    pub stmt: *const NutdbStmt,
    pub tok_type: *const u8, pub tok_start: *const u32, pub tok_end: *const u32, pub tok_kw: *const u8,
    pub node: *const NutdbNode,      // expanded nodes: NULL in batches the library produces
    pub err: *const NutdbError,
    pub impl_: *mut c_void,
    pub pnode: *const u32,           // wire nodes, one 32-bit word each (PN_*)
    pub n_ext: u64,
    pub ext: *const NutdbNodeExt,    // side table, sorted by .index
    pub wstmt: *const u64,           // F_WIRE_STMT: status | node_count << 4 | tok_used << 34 (then `stmt` is NULL)
}
#[repr(C)]
pub struct NutdbBatchDevice {
    pub stmt: *const c_void, pub tok_type: *const c_void, pub tok_start: *const c_void, pub tok_end: *const c_void,
    pub tok_kw: *const c_void, pub node: *const c_void, pub err: *const c_void,
    pub wstmt: *const c_void,
}
#[repr(C)]
pub struct NutdbMShard {
    pub device_index: c_int, pub sql: *const u8, pub stmt_off: *const u64, pub n_stmt: u64, pub flags: u32, pub first_stmt: u64,
}
#[repr(C)]
pub struct NutdbMChunk { pub shard: u64, pub first_stmt: u64, pub device: c_int, pub on_device: c_int, pub batch: NutdbBatch }
pub type nutdb_chunk_fn = Option<unsafe extern "C" fn(user: *mut c_void, chunk: *const NutdbMChunk)>;

pub enum NutdbCtx {}
pub enum NutdbMCtx {}

pub const F_NO_TOKENS: u32 = 1;
pub const F_DEVICE_INPUT: u32 = 2;
pub const F_NO_HOST_COPY: u32 = 4;
pub const F_ALL_TOKENS: u32 = 8;
pub const F_WIRE_STMT: u32 = 16;
pub const F_OFFSETS32: u32 = 32;
pub const MF_GATHER_DEVICE0: u32 = 0x100;
pub const MF_SERIAL_CALLBACKS: u32 = 0x200;
pub const OK: c_int = 0;
pub const E_CUDA: c_int = -1;
pub const E_ARG: c_int = -2;
pub const E_NOMEM: c_int = -3;
pub const ST_OK: u32 = 0;
pub const ST_LEX_ERROR: u32 = 1;
pub const ST_SYNTAX_ERROR: u32 = 2;
pub const ST_LIMIT: u32 = 3;
pub const ST_REFERENCE_PANIC: u32 = 4;
pub const NK_FIRST_INTERIOR: u8 = 32;
// wire-node layout (NUTDB_PN_*)
pub const PN_SUB_SHIFT: u32 = 7;
pub const PN_FLAG_SHIFT: u32 = 12;
pub const PN_SIZE_SHIFT: u32 = 13;
pub const PN_SIZE_EXT: u32 = 0x7FFFF;
pub const PN_GAP_SHIFT: u32 = 13;
pub const PN_GAP_EXT: u32 = 1022;
pub const PN_GAP_NOSPAN: u32 = 1023;
pub const PN_LEN_SHIFT: u32 = 23;
pub const PN_LEN_SPECIAL: u32 = 511;

extern "C" {
    pub fn nutdb_gpu_ctx_create(device: c_int) -> *mut NutdbCtx;
    pub fn nutdb_gpu_ctx_destroy(ctx: *mut NutdbCtx);
    pub fn nutdb_gpu_last_error(ctx: *const NutdbCtx) -> *const c_char;
    pub fn nutdb_gpu_parse_batch(ctx: *mut NutdbCtx, sql: *const u8, stmt_off: *const u64, n_stmt: u64, flags: u32,
                                 out: *mut NutdbBatch) -> c_int;
    pub fn nutdb_gpu_parse(ctx: *mut NutdbCtx, sql: *const u8, len: u64, out: *mut NutdbBatch) -> c_int;
    pub fn nutdb_gpu_batch_free(ctx: *mut NutdbCtx, batch: *mut NutdbBatch);
    pub fn nutdb_batch_expand_nodes(batch: *const NutdbBatch, out: *mut NutdbNode) -> c_int;
    pub fn nutdb_batch_expand_stmts(batch: *const NutdbBatch, out: *mut NutdbStmt) -> c_int;
    pub fn nutdb_gpu_batch_device(batch: *const NutdbBatch, out: *mut NutdbBatchDevice) -> c_int;
    pub fn nutdb_gpu_batch_fetch_ext(batch: *mut NutdbBatch) -> c_int;
    pub fn nutdb_gpu_batch_hash(batch: *const NutdbBatch, out: *mut u64) -> c_int;
    pub fn nutdb_gpu_split_statements(ctx: *mut NutdbCtx, sql: *const u8, len: u64, flags: u32, stmt_off: *mut *const u64,
                                      n_stmt: *mut u64) -> c_int;
    pub fn nutdb_gpu_last_timing(ctx: *const NutdbCtx, ms: *mut f32) -> c_int;
    pub fn nutdb_gpu_last_launches(ctx: *const NutdbCtx) -> c_int;
    pub fn nutdb_gpu_set_profiling(ctx: *mut NutdbCtx, on: c_int);
    pub fn nutdb_gpu_kernel_timing(ctx: *const NutdbCtx, i: c_int, name: *mut *const c_char, ms: *mut f32) -> c_int;
    pub fn nutdb_gpu_last_slow_statements(ctx: *const NutdbCtx) -> u64;
    pub fn nutdb_gpu_last_wide_statements(ctx: *const NutdbCtx) -> u64;
    pub fn nutdb_gpu_last_exact_lexed_statements(ctx: *const NutdbCtx) -> u64;
    pub fn nutdb_gpu_ctx_stream(ctx: *const NutdbCtx) -> *mut c_void;
    pub fn nutdb_gpu_mctx_create(devices: *const c_int, n_devices: c_int, workers_per_device: c_int) -> *mut NutdbMCtx;
    pub fn nutdb_gpu_mctx_destroy(m: *mut NutdbMCtx);
    pub fn nutdb_gpu_mctx_last_error(m: *const NutdbMCtx) -> *const c_char;
    pub fn nutdb_gpu_mctx_device_count(m: *const NutdbMCtx) -> c_int;
    pub fn nutdb_gpu_mctx_parse_shards(m: *mut NutdbMCtx, shards: *const NutdbMShard, n_shards: u64, flags: u32,
                                       f: nutdb_chunk_fn, user: *mut c_void) -> c_int;
    pub fn nutdb_gpu_mctx_parse_stream(m: *mut NutdbMCtx, sql: *const u8, stmt_off: *const u64, n_stmt: u64, chunk_bytes: u64,
                                       flags: u32, f: nutdb_chunk_fn, user: *mut c_void) -> c_int;
    pub fn nutdb_gpu_copy_to_host(dst: *mut c_void, src_device: *const c_void, bytes: u64) -> c_int;
    pub fn nutdb_fmt_debug(batch: *const NutdbBatch, i: u64, sql: *const u8, len: usize, buf: *mut c_char, cap: usize) -> usize;
    pub fn nutdb_fmt_error(batch: *const NutdbBatch, i: u64, sql: *const u8, len: usize, buf: *mut c_char, cap: usize) -> usize;
    pub fn nutdb_gpu_version() -> *const c_char;
}

// node kinds (NUTDB_NK_*): leaves below NK_FIRST_INTERIOR
pub const NK_NAME: u8 = 1;
pub const NK_ALIAS: u8 = 2;
pub const NK_QUAL: u8 = 3;
pub const NK_IDENT: u8 = 4;
pub const NK_QPARAM: u8 = 5;
pub const NK_LIT_INT: u8 = 6;
pub const NK_LIT_FLOAT: u8 = 7;
pub const NK_LIT_STR: u8 = 8;
pub const NK_LIT_BOOL: u8 = 9;
pub const NK_LIT_NULL: u8 = 10;
pub const NK_LIT_INTERVAL: u8 = 11;
pub const NK_NUM: u8 = 12;
pub const NK_STR: u8 = 13;
pub const NK_DT_SCALAR: u8 = 14;
pub const NK_ORDER_DESC: u8 = 15;
pub const NK_FN_NAME: u8 = 16;
pub const NK_ENT_NAME: u8 = 17;
pub const NK_POS_FIRST: u8 = 18;
pub const NK_POS_AFTER: u8 = 19;
pub const NK_STRATEGY: u8 = 20;
pub const NK_STMT_SELECT: u8 = 32;
pub const NK_STMT_INSERT: u8 = 33;
pub const NK_ROWS: u8 = 34;
pub const NK_ROW: u8 = 35;
pub const NK_STMT_EXPLAIN: u8 = 36;
pub const NK_STMT_ALTER: u8 = 37;
pub const NK_STMT_CREATE: u8 = 38;
pub const NK_TABLEDEF: u8 = 39;
pub const NK_VIEWDEF: u8 = 40;
pub const NK_COLDEF: u8 = 41;
pub const NK_INDEXDEF: u8 = 42;
pub const NK_CONSTRDEF: u8 = 43;
pub const NK_ATTR_PK: u8 = 44;
pub const NK_ATTR_ORDER: u8 = 45;
pub const NK_ATTR_PART: u8 = 46;
pub const NK_ATTR_DEFAULT: u8 = 47;
pub const NK_STMT_DESCRIBE: u8 = 48;
pub const NK_STMT_DROP: u8 = 49;
pub const NK_STMT_TRUNCATE: u8 = 50;
pub const NK_STMT_OPTIMIZE: u8 = 51;
pub const NK_STMT_SET: u8 = 52;
pub const NK_QUERY_BODY: u8 = 53;
pub const NK_QUERY_UNION: u8 = 54;
pub const NK_WITH: u8 = 55;
pub const NK_DISTINCT: u8 = 56;
pub const NK_COLS: u8 = 57;
pub const NK_FROM: u8 = 58;
pub const NK_JOIN: u8 = 59;
pub const NK_WHERE: u8 = 60;
pub const NK_GROUPBY: u8 = 61;
pub const NK_HAVING: u8 = 62;
pub const NK_ORDERBY: u8 = 63;
pub const NK_LIMIT: u8 = 64;
pub const NK_COLLECTION: u8 = 65;
pub const NK_UNARY: u8 = 66;
pub const NK_BINARY: u8 = 67;
pub const NK_FNCALL: u8 = 68;
pub const NK_DT_PARAM: u8 = 69;
pub const NK_DT_COMPOUND: u8 = 70;
