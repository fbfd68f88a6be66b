// nutdb_gpu.cu -- CUDA kernels (sm_100a) and the C ABI of libnutdb_gpu.so (include/nutdb_gpu.h).
//
// Pipeline of one nutdb_gpu_parse_batch() call (replaces n calls of Parser::parse,
// reference src/parser/mod.rs:27):
//
//   k_prep          statement offsets -> 32-bit relative offsets + a bitmap of statement starts; validation
//   k_cuts          cut points of the lexer's ranges (statement starts nearest to the multiples of n / 2048)
//   k_lex4          the single-pass lexer (lex4_kernel.cuh, lex3_core.cuh): persistent CTAs, one range of whole
//                   statements at a time, 8 KB tiles by cp.async.bulk + mbarrier; class masks, context transition
//                   functions + scans, token masks, thread-per-token emission, keyword hash.  (k_lex3: the same for
//                   one token stream with look-back scans, the fallback for a statement too long to cut around.)
//   k_punt_list, k_lex_exact<0/1>   statements the mask lexer flagged: the exact sequential walker, one thread each
//   k_parse_fast    one thread per statement: table-driven operator-precedence parser, stack in shared memory
//   k_wide_order, k_parse_wide   the same parser for what it declined: stack in the statement's node range, wider
//                   grammar, folding; persistent warps pull groups of statements cut by a token budget
//   k_parse         what both declined: the bytecode pushdown automaton (whole grammar, every error, pull order)
//   k_parse_retry   automaton statements that overflowed its local stack (stacks in global memory)
//   k_stmt_sums, k_scan_tiles, k_finalize   dense 32-bit wire nodes / 8-byte wire statement records / error records
//   k_tok_compact, k_stmt_tok_remap         dense token arrays, only for callers that ask for tokens
//   k_lex_A..D      the chunk-parallel exact walker of the verify mode NUTDB_F_ALL_TOKENS (every token incl. white
//                   space and comments; stages its tile TRANSPOSED so 32 lanes walking 32 chunks hit 32 banks)
//   k_lex2_fn, k_split<0/1>                 the raw-buffer statement splitter (nutdb_gpu_split_statements)
//
// The host side issues all of it without looking at a device-side count, up to one synchronisation.  No CPU fallback
// anywhere: without a device every entry point fails.  The multi-GPU dispatcher (nutdb_gpu_mctx_*) is dispatch.cpp.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "lex_tables.hpp"
#include "parse_fast.cuh"
#include "parse_fast_tables.hpp"

using namespace nlex;

#define LEX_THREADS 256
#define LEX_CHUNK 32
#define LEX_TILE (LEX_THREADS * LEX_CHUNK)
#define LEX_ROW (LEX_THREADS + 1)
#define SCAN_THREADS 1024
#ifndef PARSE_THREADS
#define PARSE_THREADS 1024  // statements per CTA of the automaton: a large CTA gives its shape sort more equal statements per warp
#endif
#ifndef FAST_THREADS
#define FAST_THREADS 512
#ifndef FAST_MAX_TOKENS
#define FAST_MAX_TOKENS 64u  // statements with more tokens skip the narrow pass (config 4: narrow 2.2 -> 0.25 ms, wide unchanged)
#endif
#endif
#ifndef FAST_MINBLOCKS
#define FAST_MINBLOCKS 3
#endif
//#define FAST_THREADS_DOC     // statements per CTA in k_parse_fast (re-dealt among its warps)
#define FAST_BINS 256         // 4 statement kinds x 64 token counts
#define PARSE_STACK 160      // words of local stack in the fast path
#define FIN_THREADS 256
#define FIN_OWNER_CAP 6144u  // nodes per block of k_finalize whose owner look-up is a table (256 statements x 24 nodes)
#define NODE_SLACK 8u        // fast-path node range of a statement = its token count + NODE_SLACK
#define RETRY_NONE 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------
// device views
// ------------------------------------------------------------------------------------------
struct TileSrc {
  const uint8_t* text;
  const uint32_t* bitmap;
  const uint32_t* sm;  // transposed tile
  const uint32_t* bm;  // boundary bits of the tile, one word per chunk
  uint32_t tile_begin, n;
  __device__ __forceinline__ uint8_t byte(uint32_t p) const {
    uint32_t r = p - tile_begin;
    if (r < LEX_TILE) {
      uint32_t w = sm[((r >> 2) & 7u) * LEX_ROW + (r >> 5)];
      return (uint8_t)(w >> ((r & 3u) * 8u));
    }
    return p < n ? text[p] : (uint8_t)0;
  }
  __device__ __forceinline__ bool boundary(uint32_t p) const {
    uint32_t r = p - tile_begin;
    if (r < LEX_TILE) return (bm[r >> 5] >> (r & 31u)) & 1u;
    return p < n && ((bitmap[p >> 5] >> (p & 31u)) & 1u);
  }
};

__device__ __forceinline__ void stage_tile(const uint8_t* text, const uint32_t* bitmap, uint32_t tile_begin, uint32_t n,
                                           uint32_t* sm, uint32_t* bm) {
  const uint4* src = reinterpret_cast<const uint4*>(text + tile_begin);
#pragma unroll
  for (uint32_t k = threadIdx.x; k < LEX_TILE / 16; k += LEX_THREADS) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (tile_begin + 16u * k < n) v = __ldg(src + k);
    uint32_t chunk = k >> 1, w = (k & 1u) * 4u;
    sm[(w + 0) * LEX_ROW + chunk] = v.x;
    sm[(w + 1) * LEX_ROW + chunk] = v.y;
    sm[(w + 2) * LEX_ROW + chunk] = v.z;
    sm[(w + 3) * LEX_ROW + chunk] = v.w;
  }
  bm[threadIdx.x] = bitmap[(tile_begin >> 5) + threadIdx.x];
}

__device__ __forceinline__ void stage_tables(const LexTables* g, LexTables* s) {
  const uint32_t* a = reinterpret_cast<const uint32_t*>(g);
  uint32_t* b = reinterpret_cast<uint32_t*>(s);
  for (uint32_t i = threadIdx.x; i < sizeof(LexTables) / 4; i += blockDim.x) b[i] = a[i];
}

// ------------------------------------------------------------------------------------------
// scans
// ------------------------------------------------------------------------------------------
struct Vec8Op {
  typedef uint32_t T;
  __device__ static T identity() { return NUTDB_VEC8_ID; }
  __device__ static T then(T a, T b) { return vec8_then(a, b); }
  __device__ static T shfl_up(T v, int d) { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
};

// CSum packed in a uint4: x=count, y=nseg | escaped<<31, z=tok_start, w=stmt_start | has_tok<<31
__device__ __forceinline__ uint4 csum_pack(const CSum& c) {
  return make_uint4(c.count, c.nseg | ((uint32_t)(c.escaped != 0) << 31), c.tok_start,
                    c.stmt_start | ((uint32_t)(c.has_tok != 0) << 31));
}
__device__ __forceinline__ CSum csum_unpack(const uint4& v) {
  CSum c;
  c.count = v.x;
  c.nseg = v.y & 0x7FFFFFFFu;
  c.escaped = (uint8_t)(v.y >> 31);
  c.tok_start = v.z;
  c.stmt_start = v.w & 0x7FFFFFFFu;
  c.has_tok = (uint8_t)(v.w >> 31);
  return c;
}
struct CSumOp {
  typedef uint4 T;
  __device__ static T identity() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static T then(const T& a, const T& b) {
    T r;
    r.x = a.x + b.x;
    const bool bt = (b.w >> 31) != 0;
    const uint32_t nseg = (a.y & 0x7FFFFFFFu) + (b.y & 0x7FFFFFFFu);
    const uint32_t esc = bt ? (b.y >> 31) : ((a.y | b.y) >> 31);
    r.y = nseg | (esc << 31);
    r.z = bt ? b.z : a.z;
    const uint32_t ss = (b.y & 0x7FFFFFFFu) ? (b.w & 0x7FFFFFFFu) : (a.w & 0x7FFFFFFFu);
    r.w = ss | ((a.w | b.w) & 0x80000000u);
    return r;
  }
  __device__ static T shfl_up(const T& v, int d) {
    return make_uint4(__shfl_up_sync(0xFFFFFFFFu, v.x, d), __shfl_up_sync(0xFFFFFFFFu, v.y, d),
                      __shfl_up_sync(0xFFFFFFFFu, v.z, d), __shfl_up_sync(0xFFFFFFFFu, v.w, d));
  }
};
struct U2AddOp {
  typedef uint2 T;
  __device__ static T identity() { return make_uint2(0u, 0u); }
  __device__ static T then(const T& a, const T& b) { return make_uint2(a.x + b.x, a.y + b.y); }
  __device__ static T shfl_up(const T& v, int d) {
    return make_uint2(__shfl_up_sync(0xFFFFFFFFu, v.x, d), __shfl_up_sync(0xFFFFFFFFu, v.y, d));
  }
};

// Ordered (non-commutative) block scan.  `ws` = shared array of 32 elements.  Returns the
// exclusive prefix; `incl` = inclusive value; every thread gets the block total in `total`.
template <class Op>
__device__ __forceinline__ typename Op::T block_scan(typename Op::T v, typename Op::T* ws, typename Op::T& incl,
                                                     typename Op::T& total) {
  typedef typename Op::T T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = Op::shfl_up(v, d);
    if (lane >= d) v = Op::then(o, v);
  }
  T prev = Op::shfl_up(v, 1);
  if (lane == 0) prev = Op::identity();
  if (lane == 31) ws[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarps ? ws[lane] : Op::identity();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = Op::shfl_up(w, d);
      if (lane >= d) w = Op::then(o, w);
    }
    ws[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  T base = warp ? ws[warp - 1] : Op::identity();
  total = ws[nwarps - 1];
  incl = Op::then(base, v);
  T excl = Op::then(base, prev);
  __syncthreads();  // ws may be reused by the caller
  return excl;
}

// Single-block scan over per-tile aggregates: out[i] = exclusive prefix, *total = everything.
template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const typename Op::T* __restrict__ in,
                                                             typename Op::T* __restrict__ out, uint32_t n,
                                                             typename Op::T* __restrict__ total_out,
                                                             const uint32_t* __restrict__ n_dev = nullptr) {
  typedef typename Op::T T;
  __shared__ T ws[32];
  if (n_dev) n = min(n, *n_dev);  // the number of entries is only known on the device
  // chunks of 4 * SCAN_THREADS entries, four consecutive ones per thread: the loads of a warp are one contiguous run
  // (a thread walking its own n / SCAN_THREADS entries touched a cache line of its own per step: 0.1 ms for 48 K entries)
  T carry = Op::identity();
  for (uint32_t base = 0; base < n; base += 4u * SCAN_THREADS) {
    const uint32_t i0 = base + 4u * threadIdx.x;
    T v[4];
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) v[k] = i0 + k < n ? in[i0 + k] : Op::identity();
    T incl, total;
    T g = Op::then(carry, block_scan<Op>(Op::then(Op::then(v[0], v[1]), Op::then(v[2], v[3])), ws, incl, total));
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) {
      if (i0 + k < n) out[i0 + k] = g;
      g = Op::then(g, v[k]);
    }
    carry = Op::then(carry, total);
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// The same scans in two grid-wide passes for long inputs (one entry per 8 KB tile: 131K entries per GiB): a
// single block walking them took ~0.25 ms per scan.  Pass 1: the aggregate of each run of SCAN_THREADS entries;
// pass 2: every block reduces the aggregates before it and scans its own run.  n <= SCAN_THREADS^2.
template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan2_totals(const typename Op::T* __restrict__ in, uint32_t n,
                                                               typename Op::T* __restrict__ block_total) {
  typedef typename Op::T T;
  __shared__ T ws[32];
  const uint32_t i = blockIdx.x * SCAN_THREADS + threadIdx.x;
  T incl, total;
  block_scan<Op>(i < n ? in[i] : Op::identity(), ws, incl, total);
  if (threadIdx.x == 0) block_total[blockIdx.x] = total;
}
template <class Op>
__device__ __forceinline__ typename Op::T scan2_prefix(const typename Op::T* __restrict__ block_total, typename Op::T* ws,
                                                       typename Op::T v, typename Op::T& total_all) {
  typedef typename Op::T T;
  T incl, before;
  block_scan<Op>(threadIdx.x < blockIdx.x ? block_total[threadIdx.x] : Op::identity(), ws, incl, before);
  T total;
  const T excl = block_scan<Op>(v, ws, incl, total);
  total_all = Op::then(before, total);
  return Op::then(before, excl);
}
template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan2_apply(const typename Op::T* __restrict__ in,
                                                              typename Op::T* __restrict__ out, uint32_t n,
                                                              const typename Op::T* __restrict__ block_total,
                                                              typename Op::T* __restrict__ total_out) {
  typedef typename Op::T T;
  __shared__ T ws[32];
  const uint32_t i = blockIdx.x * SCAN_THREADS + threadIdx.x;
  T all;
  const T g = scan2_prefix<Op>(block_total, ws, i < n ? in[i] : Op::identity(), all);
  if (i < n) out[i] = g;
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = all;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan2_vec8_apply(const uint32_t* __restrict__ in,
                                                                   uint8_t* __restrict__ entry, uint32_t n,
                                                                   const uint32_t* __restrict__ block_total) {
  __shared__ uint32_t ws[32];
  const uint32_t i = blockIdx.x * SCAN_THREADS + threadIdx.x;
  uint32_t all;
  const uint32_t g = scan2_prefix<Vec8Op>(block_total, ws, i < n ? in[i] : NUTDB_VEC8_ID, all);
  if (i < n) entry[i] = (uint8_t)vec8_apply(g, 0u);
}

// The statements flagged for the exact lexer, as an ASCENDING list (the flags are raised by whichever warp sees
// the reason first, so an atomically appended list would make token positions in the extra region -- and with them
// tok_begin of those statements -- depend on scheduling).  Two passes over punt_flag around a scan of block counts.
#define PUNT_PER_THREAD 16u
#define PUNT_THREADS 256u
template <bool Scatter>
__global__ void __launch_bounds__(PUNT_THREADS) k_punt_list(const uint32_t* __restrict__ punt_flag, uint32_t nstmt,
                                                            uint2* __restrict__ block_count,
                                                            const uint2* __restrict__ block_pref,
                                                            uint32_t* __restrict__ list,
                                                            const uint32_t* __restrict__ npunt_dev) {
  __shared__ uint2 ws[32];
  if (*npunt_dev == 0u) {  // nothing flagged (the usual case): the list is empty
    if (!Scatter && threadIdx.x == 0) block_count[blockIdx.x] = make_uint2(0u, 0u);
    return;
  }
  const uint32_t s0 = (blockIdx.x * PUNT_THREADS + threadIdx.x) * PUNT_PER_THREAD;
  uint32_t bits = 0;
  for (uint32_t k = 0; k < PUNT_PER_THREAD; k++)
    if (s0 + k < nstmt && punt_flag[s0 + k] != 0u) bits |= 1u << k;
  uint2 incl, total;
  const uint2 excl = block_scan<U2AddOp>(make_uint2((uint32_t)__popc(bits), 0u), ws, incl, total);
  if (!Scatter) {
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
  } else {
    uint32_t at = block_pref[blockIdx.x].x + excl.x;
    while (bits) {
      const uint32_t k = (uint32_t)__ffs((int)bits) - 1u;
      bits &= bits - 1u;
      list[at++] = s0 + k;
    }
  }
}

// vec8 variant: stores only the ENTRY STATE of each tile (the automaton starts in state 0)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_vec8(const uint32_t* __restrict__ in, uint8_t* __restrict__ entry,
                                                            uint32_t n) {
  __shared__ uint32_t ws[32];
  const uint32_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
  uint32_t f = NUTDB_VEC8_ID;
  for (uint32_t i = lo; i < hi; i++) f = vec8_then(f, in[i]);
  uint32_t incl, total;
  uint32_t g = block_scan<Vec8Op>(f, ws, incl, total);
  for (uint32_t i = lo; i < hi; i++) {
    entry[i] = (uint8_t)vec8_apply(g, 0u);
    g = vec8_then(g, in[i]);
  }
}

// ------------------------------------------------------------------------------------------
// k_prep
// ------------------------------------------------------------------------------------------
template <typename OffT>  // uint64_t, or uint32_t with NUTDB_F_OFFSETS32
__global__ void k_prep(const OffT* __restrict__ off, uint64_t nstmt, uint64_t nbytes, uint32_t* __restrict__ off32,
                       uint32_t* __restrict__ bitmap, uint32_t* __restrict__ first_stmt, uint32_t* __restrict__ bad) {
  uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nstmt) return;
  const uint64_t base = off[0];
  const uint64_t o = off[s];
  // every offset must lie inside the batch [off[0], off[n]]: an interior offset beyond it would index the bitmap
  // (sized for the batch) out of bounds
  if (o < base || o - base > nbytes) {
    atomicOr(bad, 1u);
    off32[s] = 0u;
    return;
  }
  off32[s] = (uint32_t)(o - base);
  if (s < nstmt) {
    const uint64_t e = off[s + 1];
    if (e < o || e - base > nbytes) atomicOr(bad, 1u);
    else if (e > o) {
      uint32_t p = (uint32_t)(o - base);
      atomicOr(&bitmap[p >> 5], 1u << (p & 31u));
      atomicMin(&first_stmt[p >> 5], (uint32_t)s);
    }
  }
}

// ------------------------------------------------------------------------------------------
// lexer kernels
// ------------------------------------------------------------------------------------------
struct LexShared {
  LexTables T;
  uint32_t sm[8 * LEX_ROW];
  uint32_t bm[LEX_THREADS];
};

__global__ void __launch_bounds__(LEX_THREADS) k_lex_A(const uint8_t* __restrict__ text, const uint32_t* __restrict__ bitmap,
                                                       uint32_t n, const LexTables* __restrict__ gT,
                                                       uint32_t* __restrict__ localA, uint32_t* __restrict__ tileA) {
  __shared__ LexShared S;
  __shared__ uint32_t ws[32];
  const uint32_t tile_begin = blockIdx.x * LEX_TILE;
  stage_tables(gT, &S.T);
  stage_tile(text, bitmap, tile_begin, n, S.sm, S.bm);
  __syncthreads();
  TileSrc src{text, bitmap, S.sm, S.bm, tile_begin, n};
  const uint32_t begin = tile_begin + threadIdx.x * LEX_CHUNK;
  uint32_t f = NUTDB_VEC8_ID;
  if (begin < n) f = chunk_sim_A(S.T, src, begin, min(begin + LEX_CHUNK, n));
  uint32_t incl, total;
  uint32_t excl = block_scan<Vec8Op>(f, ws, incl, total);
  localA[blockIdx.x * LEX_THREADS + threadIdx.x] = excl;
  if (threadIdx.x == 0) tileA[blockIdx.x] = total;
}

__global__ void __launch_bounds__(LEX_THREADS) k_lex_B(const uint8_t* __restrict__ text, const uint32_t* __restrict__ bitmap,
                                                       uint32_t n, const LexTables* __restrict__ gT,
                                                       const uint32_t* __restrict__ localA,
                                                       const uint8_t* __restrict__ tileEntA, uint32_t* __restrict__ localB,
                                                       uint32_t* __restrict__ tileB) {
  __shared__ LexShared S;
  __shared__ uint32_t ws[32];
  const uint32_t tile_begin = blockIdx.x * LEX_TILE;
  stage_tables(gT, &S.T);
  stage_tile(text, bitmap, tile_begin, n, S.sm, S.bm);
  __syncthreads();
  TileSrc src{text, bitmap, S.sm, S.bm, tile_begin, n};
  const uint32_t chunk = blockIdx.x * LEX_THREADS + threadIdx.x;
  const uint32_t begin = tile_begin + threadIdx.x * LEX_CHUNK;
  uint32_t f = NUTDB_VEC8_ID;
  if (begin < n) {
    const uint8_t entA = (uint8_t)vec8_apply(localA[chunk], tileEntA[blockIdx.x]);
    f = chunk_sim_B(S.T, src, begin, min(begin + LEX_CHUNK, n), entA);
  }
  uint32_t incl, total;
  uint32_t excl = block_scan<Vec8Op>(f, ws, incl, total);
  localB[chunk] = excl;
  if (threadIdx.x == 0) tileB[blockIdx.x] = total;
}

template <bool EmitAll>
__global__ void __launch_bounds__(LEX_THREADS) k_lex_C(const uint8_t* __restrict__ text, const uint32_t* __restrict__ bitmap,
                                                       uint32_t n, const LexTables* __restrict__ gT,
                                                       const uint32_t* __restrict__ localA,
                                                       const uint8_t* __restrict__ tileEntA,
                                                       const uint32_t* __restrict__ localB,
                                                       const uint8_t* __restrict__ tileEntB,
                                                       uint4* __restrict__ localC, uint4* __restrict__ tileC) {
  __shared__ LexShared S;
  __shared__ uint4 ws[32];
  const uint32_t tile_begin = blockIdx.x * LEX_TILE;
  stage_tables(gT, &S.T);
  stage_tile(text, bitmap, tile_begin, n, S.sm, S.bm);
  __syncthreads();
  TileSrc src{text, bitmap, S.sm, S.bm, tile_begin, n};
  const uint32_t chunk = blockIdx.x * LEX_THREADS + threadIdx.x;
  const uint32_t begin = tile_begin + threadIdx.x * LEX_CHUNK;
  uint4 v = CSumOp::identity();
  if (begin < n) {
    const uint8_t entA = (uint8_t)vec8_apply(localA[chunk], tileEntA[blockIdx.x]);
    const uint8_t entB = (uint8_t)vec8_apply(localB[chunk], tileEntB[blockIdx.x]);
    v = csum_pack(chunk_count_t<EmitAll>(S.T, src, begin, begin + LEX_CHUNK, n, entA, entB));
  }
  uint4 incl, total;
  uint4 excl = block_scan<CSumOp>(v, ws, incl, total);
  localC[chunk] = excl;
  if (threadIdx.x == 0) tileC[blockIdx.x] = total;
}

struct DevSink {
  uint8_t* type;
  uint32_t* start;
  uint32_t* end;
  uint8_t* kw;
  uint32_t cap;
  uint32_t* stmt_tok_begin;
  uint32_t* stmt_tok_end;
  const uint32_t* off32;
  uint32_t nstmt;
  __device__ __forceinline__ void token(uint32_t i, uint8_t t, uint32_t s, uint32_t e, uint8_t k) {
    if (i < cap) {
      type[i] = t;
      start[i] = s;
      end[i] = e;
      kw[i] = k;
    }
  }
  // index of the (non-empty) statement starting at byte `pos`: the LAST s with off32[s] == pos
  __device__ uint32_t find_stmt(uint32_t pos) const {
    uint32_t lo = 0, hi = nstmt;  // first s in [0,nstmt) with off32[s] > pos
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (off32[mid] > pos) hi = mid;
      else lo = mid + 1;
    }
    return lo - 1;
  }
  __device__ __forceinline__ void seg_begin(uint32_t, uint32_t first, uint32_t stmt_start) {
    stmt_tok_begin[find_stmt(stmt_start)] = first;
  }
  __device__ __forceinline__ void seg_end(uint32_t, uint32_t endi, uint32_t stmt_start) {
    stmt_tok_end[find_stmt(stmt_start)] = endi;
  }
};

template <bool EmitAll>
__global__ void __launch_bounds__(LEX_THREADS) k_lex_D(const uint8_t* __restrict__ text, const uint32_t* __restrict__ bitmap,
                                                       uint32_t n, const LexTables* __restrict__ gT,
                                                       const uint32_t* __restrict__ localA,
                                                       const uint8_t* __restrict__ tileEntA,
                                                       const uint32_t* __restrict__ localB,
                                                       const uint8_t* __restrict__ tileEntB,
                                                       const uint4* __restrict__ localC,
                                                       const uint4* __restrict__ tilePrefC, DevSink sink) {
  __shared__ LexShared S;
  const uint32_t tile_begin = blockIdx.x * LEX_TILE;
  stage_tables(gT, &S.T);
  stage_tile(text, bitmap, tile_begin, n, S.sm, S.bm);
  __syncthreads();
  TileSrc src{text, bitmap, S.sm, S.bm, tile_begin, n};
  const uint32_t chunk = blockIdx.x * LEX_THREADS + threadIdx.x;
  const uint32_t begin = tile_begin + threadIdx.x * LEX_CHUNK;
  if (begin >= n) return;
  const uint8_t entA = (uint8_t)vec8_apply(localA[chunk], tileEntA[blockIdx.x]);
  const uint8_t entB = (uint8_t)vec8_apply(localB[chunk], tileEntB[blockIdx.x]);
  const CSum prefix = csum_unpack(CSumOp::then(tilePrefC[blockIdx.x], localC[chunk]));
  chunk_walk<EmitAll>(S.T, src, sink, begin, begin + LEX_CHUNK, n, entA, entB, prefix, false);
}

#include "lex2_core.cuh"
#include "lex2_kernels.cuh"
#include "lex3_kernels.cuh"
#include "lex4_kernel.cuh"

// ------------------------------------------------------------------------------------------
// parser kernels
// ------------------------------------------------------------------------------------------
struct DTok {
  const uint8_t* ty;
  const uint32_t* st;
  const uint32_t* en;
  const uint8_t* kwp;
  uint32_t n;
  __device__ __forceinline__ uint8_t type(uint32_t i) const { return i < n ? ty[i] : (uint8_t)NUTDB_TT_EOF; }
  // the current token of a parser never lies beyond the statement's EOF token: no bounds check
  __device__ __forceinline__ uint8_t type_at(uint32_t i) const { return ty[i]; }
  __device__ __forceinline__ uint8_t kw_at(uint32_t i) const { return kwp[i]; }
  __device__ __forceinline__ uint8_t kw(uint32_t i) const { return i < n ? kwp[i] : (uint8_t)0; }
  __device__ __forceinline__ uint32_t start(uint32_t i) const { return i < n ? st[i] : 0u; }
  __device__ __forceinline__ uint32_t end(uint32_t i) const { return i < n ? en[i] : 0u; }
  __device__ __forceinline__ uint32_t pair_at(uint32_t i) const { return (uint32_t)ty[i] | ((uint32_t)kwp[i] << 8); }
};
// The same view for k_parse_fast, whose CTA has staged the (type, keyword) pairs of its statements' tokens in
// shared memory: `rel` is the statement's first token relative to the staged window; tokens outside the window
// (a statement lexed into the extra region, or a CTA with more than FAST_TOKCAP tokens) come from global memory.
#ifndef FAST_TOKCAP
#define FAST_TOKCAP 14080u
#endif
#define FAST_DYN_SMEM (FAST_TOKCAP * 2 + FAST_STACK_DEPTH * FAST_THREADS * 8)
// k_parse_fast is tuned for FAST_MINBLOCKS resident CTAs per SM: tables + sort arrays + dynamic part + the 1 KB the
// driver reserves per CTA must fit 227 KB that many times (one CTA less costs ~25 % of the kernel's speed)
static_assert(FAST_MINBLOCKS * (sizeof(npar::FastTables) + 4 * (2 * FAST_BINS + FAST_THREADS + 16) + FAST_DYN_SMEM + 1024) <=
                  227 * 1024,
              "k_parse_fast: shared memory per CTA too large for the intended occupancy");
struct DTokS {
  const uint16_t* sm;
  uint32_t rel;
  DTok g;
  __device__ __forceinline__ uint32_t pair_at(uint32_t i) const {
    const uint32_t j = rel + i;
    return j < FAST_TOKCAP ? (uint32_t)sm[j] : g.pair_at(i);
  }
  __device__ __forceinline__ uint8_t type(uint32_t i) const { return i < g.n ? (uint8_t)pair_at(i) : (uint8_t)NUTDB_TT_EOF; }
  __device__ __forceinline__ uint8_t kw(uint32_t i) const { return i < g.n ? (uint8_t)(pair_at(i) >> 8) : (uint8_t)0; }
  __device__ __forceinline__ uint32_t start(uint32_t i) const { return g.start(i); }
  __device__ __forceinline__ uint32_t end(uint32_t i) const { return g.end(i); }
};
struct DNodes {
  uint2* p;
  uint32_t cap;
  __device__ __forceinline__ npar::CNode get(uint32_t i) const {
    uint2 v = p[i];
    npar::CNode x;
    x.kind = (uint8_t)(v.x & 0xFF);
    x.sub = (uint8_t)((v.x >> 8) & 0xFF);
    x.aux = (uint16_t)(v.x >> 16);
    x.x = v.y;
    return x;
  }
  __device__ __forceinline__ void set(uint32_t i, const npar::CNode& x) {
    p[i] = make_uint2((uint32_t)x.kind | ((uint32_t)x.sub << 8) | ((uint32_t)x.aux << 16), x.x);
  }
  __device__ __forceinline__ void set_raw(uint32_t i, uint32_t header, uint32_t x) { p[i] = make_uint2(header, x); }
  __device__ __forceinline__ uint32_t capacity() const { return cap; }
};
// First compact-node slot of statement s.  Token ranges of natively lexed statements ascend with s, so
// tok_begin + NODE_SLACK*s gives disjoint ranges of tok_count + NODE_SLACK slots; statements lexed by the
// exact walker (punt[s] = 1 + their position in the extra token region) get ranges behind all of those.
__device__ __forceinline__ size_t node_slot(uint32_t s, uint32_t tb, uint32_t nstmt, const uint32_t* __restrict__ punt) {
  const uint32_t p = punt[s];
  if (p == 0u) return (size_t)tb + (size_t)NODE_SLACK * s;
  return (size_t)NODE_SLACK * nstmt + (size_t)tb + (size_t)NODE_SLACK * (p - 1u);
}

struct DText {
  const uint8_t* p;
  uint32_t n;
  __device__ __forceinline__ uint8_t byte(uint32_t i) const { return i < n ? p[i] : (uint8_t)0; }
  __device__ __forceinline__ uint8_t raw(uint32_t i) const { return p[i]; }  // i < n is the caller's business
  // Advances p0 over four-byte words of [p0, e) that hold neither a backslash nor `quote` (1 <= p0, e <= n).  Aligned
  // word loads, one per step (a step's high word is the next one's low word), none of them beyond the statement's own
  // bytes: device input needs no padding (nutdb_gpu.h), and the text buffer itself is 16-byte aligned.
  __device__ __forceinline__ uint32_t skip_plain(uint32_t p0, uint32_t e, uint32_t quote) const {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p + p0);
    const uint32_t mis = (uint32_t)(a & 3u);
    if (p0 + 4u > e || p0 + 8u - mis > n) return p0;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a - mis);
    const uint32_t sh = mis * 8u, q4 = quote * 0x01010101u;
    uint32_t lo = __ldg(w);
    do {
      const uint32_t hi = __ldg(++w);
      const uint32_t v = __funnelshift_r(lo, hi, sh);
      const uint32_t x = v ^ 0x5C5C5C5Cu, y = v ^ q4;
      if ((((x - 0x01010101u) & ~x) | ((y - 0x01010101u) & ~y)) & 0x80808080u) break;  // a zero byte in x or y
      lo = hi;
      p0 += 4u;
    } while (p0 + 4u <= e && p0 + 8u - mis <= n);
    return p0;
  }
};

__device__ __forceinline__ void stage_parse_tables(const npar::ParseTables* g, npar::ParseTables* s) {
  const uint32_t* a = reinterpret_cast<const uint32_t*>(g);
  uint32_t* b = reinterpret_cast<uint32_t*>(s);
  for (uint32_t i = threadIdx.x; i < sizeof(npar::ParseTables) / 4; i += blockDim.x) b[i] = a[i];
}

// writes the result of one statement: its NutdbStmt and, on failure, the error record at the
// start of its node range (picked up by k_finalize)
__device__ __forceinline__ void store_result(const npar::ParseResult& res, uint32_t s, uint32_t tb, uint32_t tc,
                                             uint32_t where, const DText& tx, uint2* node_range, NutdbStmt* stmt) {
  NutdbStmt S;
  S.status = res.status;
  S.tok_begin = tb;
  S.tok_count = tc;
  S.node_begin = where;  // temporary: location of the node range (RETRY_NONE = fast-path scratch)
  S.node_count = res.status == NUTDB_ST_OK ? res.node_count : 0u;
  S.tok_used = res.tok_used;
  stmt[s] = S;
  if (res.status != NUTDB_ST_OK) {
    uint32_t line = 0, col = 0, pos = 0;
    if (res.err_has_pos) {
      pos = res.err_pos;
      npar::get_pos(tx, pos, line, col);
    }
    // the error record (32 bytes) parks in the statement's own (now unused) compact-node range
    node_range[0] = make_uint2(s, (uint32_t)res.status | ((uint32_t)res.err_code << 16));
    node_range[1] = make_uint2(line, col);
    node_range[2] = make_uint2(pos, res.err_a);
    node_range[3] = make_uint2(res.err_b, res.err_c);
  }
}

// Token range of a statement.  The single-pass lexer does not know statements: it stores, per 32-byte window, the index
// of the first token ending in it and the masks of bytes where a token / an EOF token ends; the number of tokens that
// end before byte `pos` follows from those.  Statements lexed by the exact walker have explicit ranges.
struct StmtToks {
  const uint32_t* win_idx;
  const uint32_t* win_has;
  const uint32_t* win_eof;
  const uint32_t* punt;            // per statement: 0, or 1 + position in the exact lexer's region
  const uint32_t* x_begin;         // per statement, valid where punt != 0
  const uint32_t* x_end;
  const uint32_t* cut_stmt;        // first statement of each lexer range (k_cuts); a range's windows sit `range` slots up
  const uint32_t* nranges;
  uint32_t tok_cap;                // capacity of the token arrays
  __device__ __forceinline__ uint32_t range_of(uint32_t s) const {
    uint32_t lo = 0, hi = *nranges;  // last range with cut_stmt[r] <= s
    while (hi - lo > 1u) {
      const uint32_t mid = (lo + hi) >> 1;
      if (cut_stmt[mid] <= s) lo = mid;
      else hi = mid;
    }
    return lo;
  }
  __device__ __forceinline__ uint32_t index_at(uint32_t pos, uint32_t r) const {
    const size_t w = (size_t)(pos >> 5) + r;
    const uint32_t b = pos & 31u;
    uint32_t x = win_idx[w];
    if (b) {
      const uint32_t below = (1u << b) - 1u;
      x += (uint32_t)__popc(win_has[w] & below) + (uint32_t)__popc(win_eof[w] & below);
    }
    return x;
  }
  __device__ __forceinline__ void range(uint32_t s, uint32_t o, uint32_t e, uint32_t& tb, uint32_t& tc) const {
    if (!win_idx || punt[s]) {
      tb = x_begin[s];
      tc = x_end[s] - tb;
    } else {
      const uint32_t r = range_of(s);
      tb = index_at(o, r);
      tc = index_at(e, r) - tb;
    }
    // (a batch whose token estimate was too small is run again -- but this attempt must stay inside the arrays)
    if (tb >= tok_cap || tc > tok_cap - tb) {
      tb = 0;
      tc = 0;
    }
  }
};

// Pass 1, one thread per statement: the straight-line parser (parse_fast.cuh).  Statements it
// declines go to the slow list.  Small code, no interpreter state: this is where a query log's
// bulk is parsed.
__global__ void __launch_bounds__(FAST_THREADS, FAST_MINBLOCKS) k_parse_fast(
    const uint8_t* __restrict__ text, const uint32_t* __restrict__ off32, uint32_t nstmt, uint32_t ntok,
    const uint8_t* __restrict__ tok_type, const uint32_t* __restrict__ tok_start, const uint32_t* __restrict__ tok_end,
    const uint8_t* __restrict__ tok_kw, StmtToks stoks, NutdbStmt* __restrict__ stmt, uint2* __restrict__ scratch,
    uint32_t* __restrict__ slow_list, uint32_t* __restrict__ slow_count, const uint32_t* __restrict__ punt,
    int lex_only, uint32_t tok_alloc, const npar::FastTables* __restrict__ gF, const uint32_t* __restrict__ ntok_main_dev,
    const uint32_t* __restrict__ ntok_extra_dev, const uint32_t* __restrict__ gate) {
  if (*gate) return;  // invalid statement offsets: nothing downstream of k_prep may trust them
  if (ntok_main_dev) ntok = *ntok_main_dev + *ntok_extra_dev;  // tokens of the main region + of the exact lexer's region
  // Lanes of a warp run independent parsers, so they only execute together where their statements look
  // alike.  The block therefore re-deals its statements: a counting sort in shared memory by (first keyword,
  // token-count bucket) puts statements of the same kind and similar length into the same warp.
  __shared__ uint32_t bin_count[FAST_BINS], bin_base[FAST_BINS], order[FAST_THREADS];
  __shared__ npar::FastTables FT;
  __shared__ uint32_t tok_lo, tok_hi;
  // dynamic shared memory: the staged (type, keyword) pairs of the CTA's tokens, then the operator stacks
  // (entry i of thread x at fstack[i * FAST_THREADS + x])
  extern __shared__ __align__(16) unsigned char fast_smem[];
  uint16_t* const stok = reinterpret_cast<uint16_t*>(fast_smem);
  npar::FastStackEntry* const fstack = reinterpret_cast<npar::FastStackEntry*>(fast_smem + FAST_TOKCAP * 2);
  {  // the grammar table (built on the host, parse_fast_tables.hpp)
    const uint32_t* a = reinterpret_cast<const uint32_t*>(gF);
    uint32_t* b = reinterpret_cast<uint32_t*>(&FT);
    for (uint32_t i = threadIdx.x; i < sizeof(npar::FastTables) / 4; i += FAST_THREADS) b[i] = a[i];
  }
  uint32_t s = blockIdx.x * FAST_THREADS + threadIdx.x;
  if (threadIdx.x < FAST_BINS) bin_count[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    tok_lo = 0xFFFFFFFFu;
    tok_hi = 0u;
  }
  __syncthreads();
  // The CTA's statements are consecutive, so their tokens are one contiguous run of the token arrays (except the
  // few lexed into the extra region).  Stage that run's (type, keyword) pairs in shared memory with coalesced
  // 16-byte loads: the parsers below then never wait on a global load for a token.
  uint32_t tb0 = 0xFFFFFFFFu, tc0 = 0;
  if (s < nstmt && off32[s + 1] != off32[s]) stoks.range(s, off32[s], off32[s + 1], tb0, tc0);
  {
    const uint32_t wmin = __reduce_min_sync(0xFFFFFFFFu, tb0);
    if ((threadIdx.x & 31u) == 0 && wmin != 0xFFFFFFFFu) atomicMin(&tok_lo, wmin);
  }
  __syncthreads();
  const uint32_t lo16 = tok_lo == 0xFFFFFFFFu ? 0u : (tok_lo & ~15u);
  {  // the end of the run: the furthest token of any statement that starts inside the window
    const uint32_t te0 = (tb0 != 0xFFFFFFFFu && tb0 - lo16 < FAST_TOKCAP) ? tb0 + tc0 : 0u;
    const uint32_t wmax = __reduce_max_sync(0xFFFFFFFFu, te0);
    if ((threadIdx.x & 31u) == 0 && wmax != 0u) atomicMax(&tok_hi, wmax);
  }
  __syncthreads();
  if (!lex_only) {
    const uint32_t hi = min(max(tok_hi, lo16), lo16 + FAST_TOKCAP);
    for (uint32_t v = threadIdx.x; v * 16u < hi - lo16; v += FAST_THREADS) {
      const uint32_t base = lo16 + v * 16u;
      if (base + 16u > tok_alloc) break;
      const uint4 a = *reinterpret_cast<const uint4*>(tok_type + base);
      const uint4 b = *reinterpret_cast<const uint4*>(tok_kw + base);
      uint4 x, y;
      x.x = __byte_perm(a.x, b.x, 0x5140);
      x.y = __byte_perm(a.x, b.x, 0x7362);
      x.z = __byte_perm(a.y, b.y, 0x5140);
      x.w = __byte_perm(a.y, b.y, 0x7362);
      y.x = __byte_perm(a.z, b.z, 0x5140);
      y.y = __byte_perm(a.z, b.z, 0x7362);
      y.z = __byte_perm(a.w, b.w, 0x5140);
      y.w = __byte_perm(a.w, b.w, 0x7362);
      uint4* d = reinterpret_cast<uint4*>(stok + v * 16u);
      d[0] = x;
      d[1] = y;
    }
  }
  __syncthreads();
  uint32_t key = FAST_BINS - 1, rank = 0;
  if (tb0 != 0xFFFFFFFFu) {
    const uint32_t j0 = tb0 - lo16;
    const uint32_t p0 = (j0 < FAST_TOKCAP && !lex_only) ? (uint32_t)stok[j0] : ((uint32_t)tok_type[tb0] | ((uint32_t)tok_kw[tb0] << 8));
    const uint32_t k0 = (p0 & 0xFFu) == NUTDB_TT_KeywordOrIdentifier ? (p0 >> 8) : 0u;
    const uint32_t kind = k0 == npar::KW_SELECT ? 0u : (k0 == npar::KW_INSERT ? 1u : (k0 == npar::KW_CREATE ? 2u : 3u));
    key = kind * 64u + min(63u, tc0);  // equal token counts in a warp: its lanes finish together (a bit of token-type
                                       // hash in the key instead of length resolution measured 1 % slower)
  }
  rank = atomicAdd(&bin_count[key], 1u);
  __syncthreads();
  {  // exclusive prefix over the bins (first FAST_BINS threads: warp scans + the warp totals)
    __shared__ uint32_t bin_wtot[FAST_BINS / 32];
    uint32_t v = 0, incl = 0;
    if (threadIdx.x < FAST_BINS) {
      v = bin_count[threadIdx.x];
      incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o2 = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((threadIdx.x & 31u) >= (uint32_t)d) incl += o2;
      }
      if ((threadIdx.x & 31u) == 31u) bin_wtot[threadIdx.x >> 5] = incl;
    }
    __syncthreads();
    if (threadIdx.x < FAST_BINS) {
      uint32_t before = 0;
      for (uint32_t q = 0; q < (threadIdx.x >> 5); q++) before += bin_wtot[q];
      bin_base[threadIdx.x] = before + incl - v;
    }
  }
  __syncthreads();
  order[bin_base[key] + rank] = threadIdx.x;
  __syncthreads();
  static_assert(FAST_THREADS == 2 * FAST_BINS, "the two bin arrays hold one packed range per thread");
  // the token ranges found above travel with the statements (the bin arrays are free now): the second look-up -- a
  // binary search of the range table plus four window loads -- is only repeated for the rare range that does not pack
  {
    uint32_t packed = 0xFFFFFFFFu;
    if (tb0 != 0xFFFFFFFFu && tb0 >= lo16 && tb0 - lo16 < (1u << 18) && tc0 < (1u << 14)) packed = ((tb0 - lo16) << 14) | tc0;
    (threadIdx.x < FAST_BINS ? bin_count[threadIdx.x] : bin_base[threadIdx.x - FAST_BINS]) = packed;
  }
  __syncthreads();
  const uint32_t src_thread = order[threadIdx.x];
  const uint32_t packed_range = src_thread < FAST_BINS ? bin_count[src_thread] : bin_base[src_thread - FAST_BINS];
  s = blockIdx.x * FAST_THREADS + src_thread;
  if (s >= nstmt) return;
  const uint32_t o = off32[s], len = off32[s + 1] - o;
  if (len == 0) {
    // Parser::parse(""): the first token is EOF => EmptyQuery (mod.rs:141-144); k_finalize writes the record
    NutdbStmt S;
    S.status = lex_only ? NUTDB_ST_OK : NUTDB_ST_SYNTAX_ERROR;
    S.tok_begin = ntok;
    S.tok_count = 0;
    S.node_begin = RETRY_NONE;
    S.node_count = 0;
    S.tok_used = lex_only ? 0 : 1;
    stmt[s] = S;
    return;
  }
  uint32_t tb, tc;
  if (packed_range != 0xFFFFFFFFu) {
    tb = lo16 + (packed_range >> 14);
    tc = packed_range & 0x3FFFu;
  } else {
    stoks.range(s, o, o + len, tb, tc);
  }
  if (lex_only) {
    NutdbStmt S;
    S.status = NUTDB_ST_OK;
    S.tok_begin = tb;
    S.tok_count = tc;
    S.node_begin = RETRY_NONE;
    S.node_count = 0;
    S.tok_used = 0;
    stmt[s] = S;
    return;
  }
  DTokS tk{stok, tb - lo16, DTok{tok_type + tb, tok_start + tb, tok_end + tb, tok_kw + tb, tc}};
  uint2* range = scratch + node_slot(s, tb, nstmt, punt);
  DNodes nd{range, tc + NODE_SLACK};
  DText tx{text + o, len};
  npar::ParseResult res;
  npar::FastParser<DTokS, DNodes, DText> f(&FT, tk, nd, tx, fstack + threadIdx.x, FAST_THREADS);
  // (a long statement goes straight to the wide pass: its tokens overflow this CTA's staged window, it nests deeper
  // than eight entries more often than not, and its lane would hold the warp long after the others are done)
  if (tc <= FAST_MAX_TOKENS && f.try_parse(res)) {
    store_result(res, s, tb, tc, RETRY_NONE, tx, range, stmt);
  } else {
    NutdbStmt S;  // token range for the slow pass
    S.status = NUTDB_ST_LIMIT;
    S.tok_begin = tb;
    S.tok_count = tc;
    S.node_begin = RETRY_NONE;
    S.node_count = 0;
    S.tok_used = 0;
    stmt[s] = S;
    slow_list[atomicAdd(slow_count, 1u)] = s;
  }
}

// Pass 1b, one thread per statement the first pass declined: the WIDE instantiation of the table-driven parser
// (parse_fast.cuh) -- array / map literals, index access, prefix ~, IF .. END, parenthesised subqueries, and any nesting
// depth: its operator stack is the top end of the statement's own compact-node range (the nodes grow from the bottom),
// so a deep statement needs neither a local-memory stack nor a second run.  Tokens come straight from global memory
// (these statements are long: a lane walks its own run of the arrays, the L1 serves it).  What it declines too goes to
// the exact automaton.
//
// These statements differ in length by orders of magnitude, and their lanes rarely run in step (different statements
// are in different grammar states: 5.5 of 32 lanes active on config 4), so a warp costs close to the SUM of its lanes'
// work and the kernel used to last as long as the one warp that drew 32 of the longest statements.  Hence:
// (1) k_wide_order sorts every WIDE_POOL consecutive statements of the list by length and shape and cuts the sorted run
//     into GROUPS of at most 32 statements and at most WIDE_BUDGET tokens -- a 2000-token statement gets a warp to
//     itself, short ones still share one -- listed heavy groups first;
// (2) k_parse_wide is PERSISTENT: its warps pull group after group and never wait for one another -- the heavy
//     groups from one global queue, the others pool-wise per CTA (see the kernel).
#ifndef WIDE_THREADS
#define WIDE_THREADS 256
#endif
#ifndef WIDE_POOL
#define WIDE_POOL 512  // statements sorted together, and the unit a CTA works through (config 4: 128 -> 4.7 ms, 256 -> 4.4, 512 -> 3.8, 1024 -> 3.9 + a slower sort)
#endif
#ifndef WIDE_CTAS_PER_SM
#define WIDE_CTAS_PER_SM 3  // (70 registers: three 256-thread CTAs fit an SM; capped at 64 / 48 registers for four / five
                            // CTAs the uniform config 1 got 25 % slower -- more lanes thrash the L1 -- and config 4 no faster)
#endif
#ifndef WIDE_BUDGET
#define WIDE_BUDGET 8192u  // tokens per group (config 4, global queue: 1024 -> 7.2 ms, 2048 -> 6.0, 4096 -> 4.7, 6144 / 8192 -> 4.5,
                           // 16384 -> 5.5; pool-wise: 4096 -> 4.0, 6144 -> 3.85, 8192 -> 3.7, 12288 -> 4.1)
#endif
#ifndef WIDE_MINBLOCKS
#define WIDE_MINBLOCKS 1
#endif
#ifndef WIDE_HEAVY
#define WIDE_HEAVY 512u    // a group with a statement of this many tokens is scheduled before the others
#endif
// counters[0] = heavy groups (listed from the front of `groups`), counters[1] = the others (listed from its back, the
// groups of one pool next to each other: `pools` says where)
__global__ void __launch_bounds__(WIDE_POOL) k_wide_order(const uint32_t* __restrict__ slow_list, const uint32_t* __restrict__ nslow_dev,
                                                          const uint8_t* __restrict__ tok_type, const NutdbStmt* __restrict__ stmt,
                                                          uint32_t* __restrict__ order, uint2* __restrict__ groups, uint32_t group_cap,
                                                          uint32_t* __restrict__ counters, uint2* __restrict__ pools,
                                                          uint32_t* __restrict__ pool_count) {
  const uint32_t nslow = *nslow_dev;
  __shared__ uint32_t skey[WIDE_POOL], stc[WIDE_POOL];
  __shared__ uint2 sgrp[WIDE_POOL];
  __shared__ uint32_t sn[2], sbase[2];
  // (a grid-stride loop over the pools: the list's length is only known on the device, and it is usually short)
  for (uint32_t pool = blockIdx.x; (uint64_t)pool * WIDE_POOL < nslow; pool += gridDim.x) {
    const uint32_t i0 = pool * WIDE_POOL + threadIdx.x;
    uint32_t key = 0xFFFFFFFFu, s0 = 0xFFFFFFFFu, tc0 = 0;
    if (i0 < nslow) {
      s0 = slow_list[i0];
      const uint32_t tb0 = stmt[s0].tok_begin;
      tc0 = stmt[s0].tok_count;
      uint32_t h = 0;
      const uint32_t m = min(tc0, 12u);
      for (uint32_t q = 0; q < m; q++) h = h * 31u + tok_type[tb0 + q];
      key = (min(tc0, 0xFFFFu) << 15) | (h & 0x7FFFu);
    }
    skey[threadIdx.x] = key;
    __syncthreads();
    uint32_t rank = 0;
    for (uint32_t j = 0; j < WIDE_POOL; j++) {
      const uint32_t kj = skey[j];
      rank += (kj < key || (kj == key && j < threadIdx.x)) ? 1u : 0u;
    }
    order[pool * WIDE_POOL + rank] = s0;  // (the tail of the last pool: 0xFFFFFFFF, sorted behind everything)
    stc[rank] = tc0;
    __syncthreads();  // (everyone has read skey)
    skey[rank] = key;
    __syncthreads();
    if (threadIdx.x < 32u) {
      // Cut the sorted run, longest statement first: heavy groups collect at the front of sgrp, the others at its back.
      // Warp 0, one group per turn: lane l looks at the l-th statement below the open end, a warp scan gives the
      // budget used up to it (statements of one shape -- same length, same leading token types: a log repeats its
      // templates -- run in step and cost the warp what one of them costs), the first lane over the budget ends the group.
      const uint32_t lane = threadIdx.x;
      const uint32_t nv = min((uint32_t)WIDE_POOL, nslow - pool * WIDE_POOL);
      uint32_t nh = 0, nl = 0, hi = nv;
      while (hi > 0) {
        const bool in = lane < hi;
        const uint32_t e = hi - 1u - lane;
        uint32_t pre = 0;
        if (in) pre = (lane == 0u || skey[e] != skey[e + 1u]) ? stc[e] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, pre, d);
          if ((int)lane >= d) pre += up;
        }
        const uint32_t ok = __ballot_sync(0xFFFFFFFFu, in && (lane == 0u || pre <= WIDE_BUDGET));
        const uint32_t cnt = ok == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~ok) - 1u;  // (pre ascends: `ok` is a run of low bits)
        const uint32_t lo = hi - cnt;
        const bool heavy = stc[hi - 1u] >= WIDE_HEAVY;
        if (lane == 0u) sgrp[heavy ? nh : WIDE_POOL - 1u - nl] = make_uint2(pool * WIDE_POOL + lo, cnt);
        if (heavy) nh++;
        else nl++;
        hi = lo;
      }
      if (lane == 0u) {
        sn[0] = nh;
        sn[1] = nl;
        sbase[0] = nh ? atomicAdd(counters, nh) : 0u;
        sbase[1] = nl ? atomicAdd(counters + 1, nl) : 0u;
        if (nl) pools[atomicAdd(pool_count, 1u)] = make_uint2(sbase[1], nl);  // this pool's light groups: first, count
      }
    }
    __syncthreads();
    if (threadIdx.x < sn[0]) groups[sbase[0] + threadIdx.x] = sgrp[threadIdx.x];
    if (threadIdx.x < sn[1]) groups[group_cap - 1u - (sbase[1] + threadIdx.x)] = sgrp[WIDE_POOL - 1u - threadIdx.x];
    __syncthreads();  // (the shared arrays are reused by the next pool)
  }
}

__global__ void __launch_bounds__(WIDE_THREADS, WIDE_MINBLOCKS) k_parse_wide(
    const uint8_t* __restrict__ text, const uint32_t* __restrict__ off32, const uint32_t* __restrict__ order,
    const uint2* __restrict__ groups, uint32_t group_cap, const uint32_t* __restrict__ counters,
    const uint2* __restrict__ pools, const uint32_t* __restrict__ pool_count,
    const uint8_t* __restrict__ tok_type, const uint32_t* __restrict__ tok_start,
    const uint32_t* __restrict__ tok_end, const uint8_t* __restrict__ tok_kw, const npar::FastTables* __restrict__ gF,
    NutdbStmt* __restrict__ stmt, uint2* __restrict__ scratch, uint32_t* __restrict__ slow2_list,
    uint32_t* __restrict__ slow2_count, uint32_t nstmt, const uint32_t* __restrict__ punt, uint32_t* __restrict__ queue,
    uint32_t* __restrict__ pool_queue) {
  const uint32_t nheavy = counters[0], ngroups = nheavy + counters[1], npools = *pool_count;
  if (blockIdx.x * (WIDE_THREADS / 32u) >= ngroups) return;  // (fewer groups than resident warps: the rest of the grid has nothing to pull)
  __shared__ npar::FastTables FT;
  // the light groups: pool (index into `pools`) << 32 | next group of it; POOL_NONE before the first, POOL_DONE at the end
  __shared__ unsigned long long pool_state;
  __shared__ uint32_t pool_lock;
  constexpr uint32_t POOL_NONE = 0xFFFFFFFEu, POOL_DONE = 0xFFFFFFFFu;
  {
    const uint32_t* a = reinterpret_cast<const uint32_t*>(gF);
    uint32_t* b = reinterpret_cast<uint32_t*>(&FT);
    for (uint32_t i = threadIdx.x; i < sizeof(npar::FastTables) / 4; i += WIDE_THREADS) b[i] = a[i];
    if (threadIdx.x == 0) {
      pool_state = (unsigned long long)POOL_NONE << 32;
      pool_lock = 0u;
    }
  }
  __syncthreads();
  static_assert(sizeof(npar::FastStackEntry) == sizeof(uint2), "the wide parser's stack shares the node range");
  const uint32_t lane = threadIdx.x & 31u;
  bool heavy_phase = true;
  for (;;) {
    // ---- the next group: heavy ones from the global queue, one per trip (they decide when the kernel ends); then the
    // light ones POOL-wise per CTA -- the statements of a pool are neighbours in every array (offsets, token
    // arrays, node ranges share cache lines across statement boundaries), so the CTA's warps work through one pool's
    // groups together.  No barrier: a warp that finds the pool used up fetches the next one under a lock while the
    // others retry; warps still busy with a group of the old pool simply finish it.
    uint32_t gi = 0xFFFFFFFFu;  // index into `groups`
    if (lane == 0) {
      if (heavy_phase) {
        const uint32_t q = atomicAdd(queue, 1u);
        if (q < nheavy) gi = q;
      }
      if (gi == 0xFFFFFFFFu) {
        for (;;) {
          const unsigned long long old = atomicAdd(&pool_state, 1ull);
          const uint32_t p = (uint32_t)(old >> 32), idx = (uint32_t)old;
          if (p == POOL_DONE) break;
          if (p != POOL_NONE) {
            const uint2 d = pools[p];
            if (idx < d.y) {
              gi = group_cap - 1u - (d.x + idx);
              break;
            }
          }
          if (atomicCAS(&pool_lock, 0u, 1u) == 0u) {
            // (another warp may have moved the CTA on between this warp's ticket and the lock)
            if ((uint32_t)(*(volatile unsigned long long*)&pool_state >> 32) == p) {
              const uint32_t np = atomicAdd(pool_queue, 1u);
              atomicExch(&pool_state, (unsigned long long)(np < npools ? np : POOL_DONE) << 32);
            }
            __threadfence_block();
            atomicExch(&pool_lock, 0u);
          } else {
            __nanosleep(100);
          }
        }
      }
    }
    gi = __shfl_sync(0xFFFFFFFFu, gi, 0);
    if (gi == 0xFFFFFFFFu) break;
    heavy_phase = gi < nheavy;
    const uint2 g = groups[gi];
    if (lane < g.y) {
      const uint32_t s = order[g.x + lane];
      const uint32_t o = off32[s], len = off32[s + 1] - o;
      const uint32_t tb = stmt[s].tok_begin, tc = stmt[s].tok_count;
      DTok tk{tok_type + tb, tok_start + tb, tok_end + tb, tok_kw + tb, tc};
      uint2* range = scratch + node_slot(s, tb, nstmt, punt);
      const uint32_t cap = tc + NODE_SLACK;
      DNodes nd{range, cap};
      DText tx{text + o, len};
      npar::ParseResult res;
      npar::FastParser<DTok, DNodes, DText, true> f(&FT, tk, nd, tx, reinterpret_cast<npar::FastStackEntry*>(range) + (cap - 1), -1);
      if (f.try_parse(res)) store_result(res, s, tb, tc, RETRY_NONE, tx, range, stmt);
      else slow2_list[atomicAdd(slow2_count, 1u)] = s;
    }
    __syncwarp();
  }
}

// Pass 2, one thread per statement of the slow list: the exact bytecode automaton (parse_core.cuh)
// for the whole grammar, every error and constant folding.
__global__ void __launch_bounds__(PARSE_THREADS) k_parse(
    const uint8_t* __restrict__ text, const uint32_t* __restrict__ off32, const uint32_t* __restrict__ slow_list,
    uint32_t nslow, const uint8_t* __restrict__ tok_type, const uint32_t* __restrict__ tok_start,
    const uint32_t* __restrict__ tok_end, const uint8_t* __restrict__ tok_kw,
    const npar::ParseTables* __restrict__ gP, NutdbStmt* __restrict__ stmt, uint2* __restrict__ scratch,
    uint2* __restrict__ retry_list, uint32_t* __restrict__ retry_count, uint32_t nstmt,
    const uint32_t* __restrict__ punt, const uint32_t* __restrict__ nslow_dev) {
  if (nslow_dev) nslow = *nslow_dev;  // the slow list's length is only known on the device
  // The interpreters of a warp's lanes diverge, so a warp costs the SUM of its lanes' work.  A short list (a few long
  // statements among many the tables parsed) is therefore spread out: one statement every `spread` threads, up to one
  // statement per warp, over the blocks the launch has anyway (one thread per statement of the batch).
  uint32_t spread = 1;
  while (spread < 32u && (uint64_t)nslow * spread * 2u <= (uint64_t)gridDim.x * PARSE_THREADS) spread <<= 1;
  if ((uint64_t)blockIdx.x * PARSE_THREADS >= (uint64_t)nslow * spread) return;
  __shared__ npar::ParseTables P;
  __shared__ uint32_t skey[PARSE_THREADS], sorder[PARSE_THREADS];
  stage_parse_tables(gP, &P);
  uint32_t i;
  if (spread == 1u) {
    // Re-deal the CTA's statements so that a warp gets statements of the same SHAPE (same token count and the
    // same leading token types: query logs repeat a few templates): the lanes' interpreters then run in step.
    const uint32_t i0 = blockIdx.x * PARSE_THREADS + threadIdx.x;
    uint32_t key = 0xFFFFFFFFu;
    if (i0 < nslow) {
      const uint32_t s0 = slow_list[i0];
      const uint32_t tb0 = stmt[s0].tok_begin, tc0 = stmt[s0].tok_count;
      uint32_t h = 0;
      const uint32_t m = min(tc0, 24u);
      for (uint32_t q = 0; q < m; q++) h = h * 31u + tok_type[tb0 + q];
      key = (min(tc0, 0xFFFFu) << 15) | (h & 0x7FFFu);
    }
    skey[threadIdx.x] = key;
    __syncthreads();
    uint32_t rank = 0;
    for (uint32_t j = 0; j < PARSE_THREADS; j++) {
      const uint32_t kj = skey[j];
      rank += (kj < key || (kj == key && j < threadIdx.x)) ? 1u : 0u;
    }
    sorder[rank] = threadIdx.x;
    __syncthreads();
    i = blockIdx.x * PARSE_THREADS + sorder[threadIdx.x];
  } else {
    __syncthreads();
    const uint32_t slot = blockIdx.x * PARSE_THREADS + threadIdx.x;
    i = (slot % spread) == 0u ? slot / spread : 0xFFFFFFFFu;
  }
  if (i >= nslow) return;
  const uint32_t s = slow_list[i];
  const uint32_t o = off32[s], len = off32[s + 1] - o;
  const uint32_t tb = stmt[s].tok_begin, tc = stmt[s].tok_count;
  uint32_t stack[PARSE_STACK];
  DTok tk{tok_type + tb, tok_start + tb, tok_end + tb, tok_kw + tb, tc};
  uint2* range = scratch + node_slot(s, tb, nstmt, punt);
  DNodes nd{range, tc + NODE_SLACK};
  DText tx{text + o, len};
  npar::ParseResult res;
  npar::Machine<DTok, DNodes, DText> m(P, tk, nd, tx, stack, PARSE_STACK, res);
  m.run(NUTDB_PROGRAM_ENTRY);
  if (res.status == NUTDB_ST_LIMIT) retry_list[atomicAdd(retry_count, 1u)] = make_uint2(s, tc);
  store_result(res, s, tb, tc, RETRY_NONE, tx, range, stmt);
}

// one thread per statement of the retry list, stack and node range in global scratch
__global__ void __launch_bounds__(PARSE_THREADS) k_parse_retry(
    const uint8_t* __restrict__ text, const uint32_t* __restrict__ off32, const uint8_t* __restrict__ tok_type,
    const uint32_t* __restrict__ tok_start, const uint32_t* __restrict__ tok_end, const uint8_t* __restrict__ tok_kw,
    const npar::ParseTables* __restrict__ gP, NutdbStmt* __restrict__ stmt, const uint2* __restrict__ retry_list,
    uint32_t nretry, const uint64_t* __restrict__ node_off, const uint64_t* __restrict__ stack_off,
    uint2* __restrict__ retry_nodes, uint32_t* __restrict__ retry_stack, uint32_t spread) {
  __shared__ npar::ParseTables P;
  stage_parse_tables(gP, &P);
  __syncthreads();
  const uint32_t slot = blockIdx.x * PARSE_THREADS + threadIdx.x;  // (a short list is spread out: see k_parse)
  if (slot % spread) return;
  const uint32_t i = slot / spread;
  if (i >= nretry) return;
  const uint32_t s = retry_list[i].x;
  const uint32_t o = off32[s], len = off32[s + 1] - o;
  const NutdbStmt old = stmt[s];
  const uint32_t tb = old.tok_begin, tc = old.tok_count;
  DTok tk{tok_type + tb, tok_start + tb, tok_end + tb, tok_kw + tb, tc};
  uint2* range = retry_nodes + node_off[i];
  DNodes nd{range, (uint32_t)(node_off[i + 1] - node_off[i])};
  DText tx{text + o, len};
  npar::ParseResult res;
  npar::Machine<DTok, DNodes, DText> m(P, tk, nd, tx, retry_stack + stack_off[i],
                                       (uint32_t)(stack_off[i + 1] - stack_off[i]), res);
  m.run(NUTDB_PROGRAM_ENTRY);
  store_result(res, s, tb, tc, (uint32_t)node_off[i], tx, range, stmt);
}

__global__ void __launch_bounds__(FIN_THREADS) k_stmt_sums(const NutdbStmt* __restrict__ stmt, uint32_t nstmt,
                                                           uint2* __restrict__ tileS) {
  __shared__ uint2 ws[32];
  const uint32_t s = blockIdx.x * FIN_THREADS + threadIdx.x;
  uint2 v = make_uint2(0u, 0u);
  if (s < nstmt) {
    const uint32_t st = stmt[s].status;
    v = make_uint2(st == NUTDB_ST_OK ? stmt[s].node_count : 0u, st != NUTDB_ST_OK ? 1u : 0u);
  }
  uint2 incl, total;
  block_scan<U2AddOp>(v, ws, incl, total);
  if (threadIdx.x == 0) tileS[blockIdx.x] = total;
}

struct ExpandOut {
  uint32_t* o;  // NutdbNode records of this statement as words: [w0, parent, a, b] per node
  __device__ __forceinline__ void body(uint32_t j, uint8_t kind, uint8_t sub, uint16_t aux, uint32_t a, uint32_t b) {
    o[4 * (size_t)j] = (uint32_t)kind | ((uint32_t)sub << 8) | ((uint32_t)aux << 16);
    *reinterpret_cast<uint2*>(o + 4 * (size_t)j + 2) = make_uint2(a, b);
  }
  __device__ __forceinline__ void parent(uint32_t j, uint32_t p) { o[4 * (size_t)j + 1] = p; }
};
struct CompactSrc {
  const uint2* p;
  __device__ __forceinline__ npar::CNode operator()(uint32_t i) const {
    uint2 v = p[i];
    npar::CNode x;
    x.kind = (uint8_t)(v.x & 0xFF);
    x.sub = (uint8_t)((v.x >> 8) & 0xFF);
    x.aux = (uint16_t)(v.x >> 16);
    x.x = v.y;
    return x;
  }
};

// Dense outputs in statement order.  A block owns FIN_THREADS consecutive statements, i.e. one
// contiguous range of output nodes; one thread per output node expands the compact node (byte
// span from the token arrays, child count, parent links of its children).
__global__ void __launch_bounds__(FIN_THREADS) k_finalize(NutdbStmt* __restrict__ stmt, uint32_t nstmt,
                                                          const uint2* __restrict__ tilePref,
                                                          const uint2* __restrict__ scratch,
                                                          const uint2* __restrict__ retry_nodes,
                                                          const uint32_t* __restrict__ tok_start,
                                                          const uint32_t* __restrict__ tok_end,
                                                          const uint32_t* __restrict__ punt,
                                                          uint32_t* __restrict__ node_out, uint4* __restrict__ err_out,
                                                          uint32_t* __restrict__ ext_count, uint4* __restrict__ ext_out,
                                                          uint32_t ext_cap, unsigned long long* __restrict__ wstmt,
                                                          uint32_t* __restrict__ wstmt_overflow) {
  __shared__ uint2 ws[32];
  __shared__ uint32_t lbegin[FIN_THREADS + 1];
  __shared__ uint32_t ltok[FIN_THREADS];
  __shared__ const uint2* lsrc[FIN_THREADS];
  const uint32_t s = blockIdx.x * FIN_THREADS + threadIdx.x;
  uint2 v = make_uint2(0u, 0u);
  const uint2* src = nullptr;
  NutdbStmt S;
  S.status = NUTDB_ST_OK;
  S.tok_begin = 0;
  if (s < nstmt) {
    S = stmt[s];
    v = make_uint2(S.status == NUTDB_ST_OK ? S.node_count : 0u, S.status != NUTDB_ST_OK ? 1u : 0u);
    src = S.node_begin == RETRY_NONE ? (S.tok_count ? scratch + node_slot(s, S.tok_begin, nstmt, punt) : scratch)
                                     : retry_nodes + S.node_begin;
  }
  uint2 incl, total;
  const uint2 excl = block_scan<U2AddOp>(v, ws, incl, total);
  const uint2 base = tilePref[blockIdx.x];
  lbegin[threadIdx.x] = excl.x;
  lsrc[threadIdx.x] = src;
  ltok[threadIdx.x] = S.tok_begin;
  if (threadIdx.x == 0) lbegin[FIN_THREADS] = total.x;
  if (s < nstmt) {
    stmt[s].node_begin = base.x + excl.x;
    if (wstmt) {  // the 8-byte wire form of the record (NUTDB_F_WIRE_STMT): status | node count << 4 | tokens pulled << 34
      const uint32_t nc = S.status == NUTDB_ST_OK ? S.node_count : 0u;
      if (nc >= (1u << 30) || S.tok_used >= (1u << 30)) atomicOr(wstmt_overflow, 1u);
      wstmt[s] = (unsigned long long)(S.status & 15u) | ((unsigned long long)(nc & 0x3FFFFFFFu) << 4) |
                 ((unsigned long long)(S.tok_used & 0x3FFFFFFFu) << 34);
    }
    if (v.y) {
      uint4 e0, e1;
      if (S.tok_count == 0) {  // empty statement: EmptyQuery, no position
        e0 = make_uint4(s, (uint32_t)NUTDB_ST_SYNTAX_ERROR | ((uint32_t)NUTDB_SE_EmptyQuery << 16), 0u, 0u);
        e1 = make_uint4(0u, 0u, 0u, 0u);
      } else {
        const uint2 r0 = src[0], r1 = src[1], r2 = src[2], r3 = src[3];
        e0 = make_uint4(r0.x, r0.y, r1.x, r1.y);
        e1 = make_uint4(r2.x, r2.y, r3.x, r3.y);
      }
      err_out[2 * (size_t)(base.y + excl.y)] = e0;
      err_out[2 * (size_t)(base.y + excl.y) + 1] = e1;
    }
  }
  __syncthreads();
  const uint32_t nblock = lbegin[FIN_THREADS];
  // node -> owning statement.  The usual block (a few thousand nodes) gets a byte table filled by the statements'
  // own threads; a block with long statements falls back to a binary search per node.
  __shared__ uint8_t owner[FIN_OWNER_CAP];
  static_assert(FIN_THREADS <= 256, "owner[] holds statement indices in a byte");
  const bool use_table = nblock <= FIN_OWNER_CAP;
  if (use_table && v.x) {
    for (uint32_t q = excl.x; q < excl.x + v.x; q++) owner[q] = (uint8_t)threadIdx.x;
  }
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < nblock; j += FIN_THREADS) {
    uint32_t lo;
    if (use_table) {
      lo = owner[j];
    } else {
      uint32_t hi = FIN_THREADS;  // last k with lbegin[k] <= j
      lo = 0;
      while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (lbegin[mid] <= j) lo = mid;
        else hi = mid;
      }
    }
    // `lo` owns node j (statements without nodes share their offset with the next one and are skipped
    // by "last k"); its nodes are [lbegin[lo], lbegin[lo+1]) of this block's range.
    // The WIRE node: one 32-bit word (nutdb_gpu.h, NUTDB_PN_*).  An interior node carries its subtree SIZE; a leaf
    // carries its byte span as (gap to the end of the leaf before it, length) -- leaves come in text order, so the
    // reader keeps a running position.  Whatever does not fit (a long literal, a long comment in front of a token, a
    // huge subtree) goes to the side table with its exact fields.
    const uint2* const nodes_s = lsrc[lo];
    const uint32_t jj = j - lbegin[lo];
    const uint2 c = nodes_s[jj];
    const uint32_t kind = c.x & 0xFFu, sub = (c.x >> 8) & 0xFFu, aux = c.x >> 16;
    uint32_t w = kind | ((sub & 31u) << NUTDB_PN_SUB_SHIFT) | ((aux & 1u) << NUTDB_PN_FLAG_SHIFT);
    bool esc = sub > 31u || kind > 127u;
    uint32_t ea = 0, eb = 0;
    if (kind < NUTDB_NK_FIRST_INTERIOR) {
      if (c.y == NUTDB_CN_NOTOK) {
        w |= (NUTDB_PN_GAP_NOSPAN << NUTDB_PN_GAP_SHIFT) | (NUTDB_PN_LEN_SPECIAL << NUTDB_PN_LEN_SHIFT);
        esc = esc || aux > 1u;
      } else {
        const uint32_t a = tok_start[ltok[lo] + c.y], b = tok_end[ltok[lo] + c.y];
        uint32_t prev_end = 0;  // end of the nearest leaf with a span in front of this one
        for (uint32_t k = jj; k-- > 0;) {
          const uint2 p = nodes_s[k];
          if ((p.x & 0xFFu) < NUTDB_NK_FIRST_INTERIOR && p.y != NUTDB_CN_NOTOK) {
            prev_end = tok_end[ltok[lo] + p.y];
            break;
          }
        }
        ea = a;
        eb = b - a;
        esc = esc || aux > 1u || a < prev_end || a - prev_end > NUTDB_PN_GAP_MAX || b - a > NUTDB_PN_LEN_MAX;
        w |= ((a - prev_end) << NUTDB_PN_GAP_SHIFT) | ((b - a) << NUTDB_PN_LEN_SHIFT);  // (overwritten below when escaped)
      }
      if (esc) w = (w & ((1u << NUTDB_PN_GAP_SHIFT) - 1u)) | (NUTDB_PN_GAP_EXT << NUTDB_PN_GAP_SHIFT) | (NUTDB_PN_LEN_SPECIAL << NUTDB_PN_LEN_SHIFT);
    } else {
      const uint32_t size = jj - c.y;  // (c.y = subtree start, statement relative)
      ea = c.y;
      esc = esc || aux > 1u || size > NUTDB_PN_SIZE_MAX;
      w |= (esc ? NUTDB_PN_SIZE_EXT : size) << NUTDB_PN_SIZE_SHIFT;
    }
    if (esc) {
      const uint32_t q = atomicAdd(ext_count, 1u);
      if (q < ext_cap) ext_out[q] = make_uint4(base.x + j, c.x, ea, eb);
    }
    node_out[(size_t)base.x + j] = w;
  }
}

// ------------------------------------------------------------------------------------------
// host side: context, buffers, the C ABI
// ------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct NutdbCtx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {};
  std::string err;
  LexTables* dLex = nullptr;
  npar::ParseTables* dPar = nullptr;
  npar::FastTables* dFast = nullptr;
  cudaStream_t stream2 = nullptr;  // the exact-lexer chain runs here beside the emit pass
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  nlex2::Lex2Tables* dLex2 = nullptr;
  uint32_t epoch = 0;       // launch number of k_lex3 (tags its look-back descriptors)
  int sm_count = 148;
  size_t tok_cap_min = 0;   // token capacity a previous batch turned out to need
  uint32_t tok_cap_num = 4; // segment capacity of k_lex4 in sixteenths of a token per byte (doubles on overflow)
  bool force_lookback = false, last_lookback = false;
  uint32_t n_punt = 0;  // statements of the last batch lexed by the exact walker
  // device buffers (grow only)
  bool debug_tiles = false, debug_timing = false;
  DevBuf dbgTim;
  uint32_t dbg_ntiles = 0;
  bool debug_sync = false;  // NUTDB_GPU_DEBUG_SYNC=1: synchronise after every launch and name the kernel that failed
  DevBuf dbgTiles;
  DevBuf extNodes;
  HostBuf hExt;
  std::vector<NutdbNodeExt> ext;
  uint32_t ext_count = 0;  // side-table entries of the last batch (on the device until fetched)
  DevBuf rangeByte, rangeStmt, rangeTokBase, rangeCount, rangeDense, tokTypeD, tokStartD, tokEndD, tokKwD;
  DevBuf winIdx, winHas, winEof, descFn, descA, descI, descB, descC, winCtx, winFn, scanTotals, hashAcc, puntBlockCount, puntBlockPref, text, off64, off32, bitmap, localA, localB, localC, tileA, tileB, tileC, tilePrefC, entA, entB, tokType, tokStart,
      tokEnd, tokKw, stmtTokBegin, stmtTokEnd, stmt, scratch, retryList, retryNodeOff, retryStackOff, retryNodes, retryStack,
      tileS, tilePrefS, nodes, errs, small, slowList, slowList2, wideOrder, wideGroups, widePools, wstmt, puntFlag, puntList, puntCounts, puntOffs, winCount, winMasks, firstStmt, winState, splitLocal, splitTile, splitPref, splitOff;
  // pinned host buffers
  HostBuf hSmall, hStmt, hTokType, hTokStart, hTokEnd, hTokKw, hNode, hErr, hRetry, hSplit;
  float ms[5] = {0, 0, 0, 0, 0};
  int launches = 0;
  uint32_t n_slow = 0;  // statements of the last batch that needed the exact automaton
  uint32_t n_wide = 0;  // ... that the wide table-driven pass parsed
  // optional per-kernel timing (nutdb_gpu_set_profiling): events around every launch
  bool profiling = false;
  struct KRec {
    const char* name;
    cudaEvent_t a, b;
  };
  std::vector<KRec> recs;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<std::pair<const char*, float>> kernel_ms;
  cudaEvent_t take_event() {
    if (ev_used == ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev_pool.push_back(e);
    }
    return ev_pool[ev_used++];
  }
  bool batch_live = false;
  NutdbBatchDevice dev_view{};
};

namespace {

bool ck(NutdbCtx* c, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  c->err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}
#define CK(call)                         \
  do {                                   \
    if (!ck(ctx, (call), #call)) return NUTDB_E_CUDA; \
  } while (0)

int ensure_dev(NutdbCtx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return NUTDB_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? NUTDB_E_NOMEM : NUTDB_E_CUDA;
  }
  b.cap = want;
  return NUTDB_OK;
}
// grow-only buffer whose NEW storage starts zeroed (look-back descriptors: stale epochs must not look current)
int ensure_dev_zeroed(NutdbCtx* ctx, DevBuf& b, size_t bytes, cudaStream_t st) {
  if (bytes <= b.cap) return NUTDB_OK;
  const int rc = ensure_dev(ctx, b, bytes);
  if (rc != NUTDB_OK) return rc;
  return cudaMemsetAsync(b.p, 0, b.cap, st) == cudaSuccess ? NUTDB_OK : NUTDB_E_CUDA;
}
int ensure_host(NutdbCtx* ctx, HostBuf& b, size_t bytes) {
  if (bytes <= b.cap) return NUTDB_OK;
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMallocHost(&b.p, want);
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return NUTDB_E_NOMEM;
  }
  b.cap = want;
  return NUTDB_OK;
}
// LAUNCH(name, kernel<<<...>>>(...)): counts the launch and, in profiling mode, brackets it with events
#define LAUNCH_ON(stream_, name, ...)                      \
  do {                                                     \
    NutdbCtx::KRec r_{name, nullptr, nullptr};             \
    if (ctx->profiling) {                                  \
      r_.a = ctx->take_event();                            \
      r_.b = ctx->take_event();                            \
      cudaEventRecord(r_.a, stream_);                      \
    }                                                      \
    __VA_ARGS__;                                           \
    ctx->launches++;                                       \
    if (ctx->debug_sync) {                                 \
      const cudaError_t e_ = cudaStreamSynchronize(stream_); \
      if (e_ != cudaSuccess) {                             \
        ctx->err = std::string(name) + ": " + cudaGetErrorString(e_); \
        fprintf(stderr, "nutdb_gpu: %s failed: %s\n", name, cudaGetErrorString(e_)); \
        return NUTDB_E_CUDA;                               \
      }                                                    \
    }                                                      \
    if (ctx->profiling) {                                  \
      cudaEventRecord(r_.b, stream_);                      \
      ctx->recs.push_back(r_);                             \
    }                                                      \
  } while (0)
#define LAUNCH(name, ...) LAUNCH_ON(st, name, __VA_ARGS__)
#define ENSURE_DEV(buf, bytes)                      \
  do {                                              \
    int rc_ = ensure_dev(ctx, ctx->buf, (bytes));   \
    if (rc_ != NUTDB_OK) return rc_;                \
  } while (0)
#define ENSURE_HOST(buf, bytes)                     \
  do {                                              \
    int rc_ = ensure_host(ctx, ctx->buf, (bytes));  \
    if (rc_ != NUTDB_OK) return rc_;                \
  } while (0)

// the side table of the last batch's wire nodes: device -> host, sorted by node index
int fetch_ext(NutdbCtx* ctx) {
  const uint32_t nl = ctx->ext_count;
  ctx->ext.clear();
  if (!nl) return NUTDB_OK;
  ENSURE_HOST(hExt, 16 * (size_t)nl);
  CK(cudaMemcpy(ctx->hExt.p, ctx->extNodes.p, 16 * (size_t)nl, cudaMemcpyDeviceToHost));
  const NutdbNodeExt* hp = (const NutdbNodeExt*)ctx->hExt.p;
  ctx->ext.assign(hp, hp + nl);
  std::sort(ctx->ext.begin(), ctx->ext.end(), [](const NutdbNodeExt& x, const NutdbNodeExt& y) { return x.index < y.index; });
  return NUTDB_OK;
}

void free_all(NutdbCtx* c) {
  DevBuf* d[] = {&c->extNodes, &c->rangeByte, &c->rangeStmt, &c->rangeTokBase, &c->rangeCount, &c->rangeDense, &c->tokTypeD, &c->tokStartD, &c->tokEndD, &c->tokKwD, &c->dbgTim, &c->dbgTiles, &c->winIdx, &c->winHas, &c->winEof, &c->descFn, &c->descA, &c->descI, &c->descB, &c->descC, &c->winCtx, &c->winFn, &c->scanTotals, &c->hashAcc, &c->puntBlockCount, &c->puntBlockPref, &c->text, &c->off64, &c->off32, &c->bitmap, &c->localA, &c->localB, &c->localC, &c->tileA, &c->tileB,
                 &c->tileC, &c->tilePrefC, &c->entA, &c->entB, &c->tokType, &c->tokStart, &c->tokEnd, &c->tokKw,
                 &c->stmtTokBegin, &c->stmtTokEnd, &c->stmt, &c->scratch, &c->retryList, &c->retryNodeOff,
                 &c->retryStackOff, &c->retryNodes, &c->retryStack, &c->tileS, &c->tilePrefS, &c->nodes, &c->errs,
                 &c->small, &c->slowList, &c->slowList2, &c->wideOrder, &c->wideGroups, &c->widePools, &c->wstmt, &c->puntFlag, &c->puntList, &c->puntCounts, &c->puntOffs, &c->winCount, &c->winMasks, &c->firstStmt, &c->winState, &c->splitLocal,
                 &c->splitTile, &c->splitPref, &c->splitOff};
  for (DevBuf* b : d)
    if (b->p) cudaFree(b->p);
  HostBuf* h[] = {&c->hExt, &c->hSmall, &c->hStmt, &c->hTokType, &c->hTokStart, &c->hTokEnd, &c->hTokKw, &c->hNode, &c->hErr,
                  &c->hRetry, &c->hSplit};
  for (HostBuf* b : h)
    if (b->p) cudaFreeHost(b->p);
}

}  // namespace

extern "C" {

const char* nutdb_gpu_version(void) { return "nutdb-gpu 0.1 (sm_100a)"; }

// A stream whose grids are dispatched ahead of the main stream's: the exact-lexer chain is a series of tiny dependent
// kernels, and at equal priority each of them would queue behind all 131K blocks of the emit pass it runs beside.
static bool create_priority_stream(cudaStream_t* s) {
  int least = 0, greatest = 0;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return false;
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, greatest) == cudaSuccess;
}

NutdbCtx* nutdb_gpu_ctx_create(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return nullptr;  // no CPU fallback: without a CUDA device there is no context
  }
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  NutdbCtx* ctx = new (std::nothrow) NutdbCtx();
  if (!ctx) return nullptr;
  ctx->device = device;
  bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
            create_priority_stream(&ctx->stream2) &&
            cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 6 && ok; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  LexTables lt;
  try {
    build_lex_tables(lt);
  } catch (...) {
    ok = false;
  }
  ok = ok && cudaMalloc(&ctx->dLex, sizeof(LexTables)) == cudaSuccess &&
       cudaMalloc(&ctx->dPar, sizeof(npar::ParseTables)) == cudaSuccess &&
       cudaMemcpy(ctx->dLex, &lt, sizeof(lt), cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(ctx->dPar, &npar::PARSE_TABLES, sizeof(npar::ParseTables), cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) {
    static_assert(sizeof(npar::FastTables) % 4 == 0, "FastTables is staged word by word");
    npar::FastTables ft;
    try {
      npar::fast_tables_build(ft);
    } catch (...) {
      ok = false;
    }
    ok = ok && cudaMalloc(&ctx->dFast, sizeof(ft)) == cudaSuccess &&
         cudaMemcpy(ctx->dFast, &ft, sizeof(ft), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (ok) {
    nlex2::Lex2Tables l2;
    nlex2::build_lex2_tables(l2);
    ok = cudaMalloc(&ctx->dLex2, sizeof(l2)) == cudaSuccess &&
         cudaMemcpy(ctx->dLex2, &l2, sizeof(l2), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (ok) ok = ensure_dev(ctx, ctx->small, 256) == NUTDB_OK && ensure_host(ctx, ctx->hSmall, 256) == NUTDB_OK;
  // k_parse_fast: staged tokens (static) + operator stacks (dynamic) exceed the 48 KB default
  if (ok) ok = cudaFuncSetAttribute(k_parse_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_DYN_SMEM) == cudaSuccess;
  if (ok) ok = cudaFuncSetAttribute(k_lex3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Lex3Shared)) == cudaSuccess;
  if (ok) ok = cudaFuncSetAttribute(k_lex4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Lex4Shared)) == cudaSuccess;
  if (ok) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->sm_count = sms;
    if (const char* e = std::getenv("NUTDB_GPU_DEBUG_SYNC")) ctx->debug_sync = e[0] == '1';
  }
  if (!ok) {
    nutdb_gpu_ctx_destroy(ctx);
    return nullptr;
  }
  return ctx;
}

void nutdb_gpu_ctx_destroy(NutdbCtx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  free_all(ctx);
  if (ctx->dLex) cudaFree(ctx->dLex);
  if (ctx->dPar) cudaFree(ctx->dPar);
  if (ctx->dFast) cudaFree(ctx->dFast);
  if (ctx->dLex2) cudaFree(ctx->dLex2);
  for (int i = 0; i < 6; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->evFork) cudaEventDestroy(ctx->evFork);
  if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* nutdb_gpu_last_error(const NutdbCtx* ctx) { return ctx ? ctx->err.c_str() : "no context (no CUDA device?)"; }

int nutdb_gpu_parse_batch(NutdbCtx* ctx, const uint8_t* sql, const uint64_t* stmt_off, uint64_t n_stmt, uint32_t flags,
                          NutdbBatch* out) {
  if (!ctx) return NUTDB_E_CUDA;
  if (!out || (!stmt_off) || (n_stmt > 0 && !sql && !(flags & NUTDB_F_DEVICE_INPUT))) {
    ctx->err = "null argument";
    return NUTDB_E_ARG;
  }
  if (n_stmt >= 0x7FFFFFFFull) {
    ctx->err = "too many statements in one batch";
    return NUTDB_E_ARG;
  }
  std::memset(out, 0, sizeof(*out));
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const bool dev_in = (flags & NUTDB_F_DEVICE_INPUT) != 0;
  const bool lex_only = (flags & NUTDB_F_ALL_TOKENS) != 0;
  const uint32_t nstmt = (uint32_t)n_stmt;
  ctx->launches = 0;
  ctx->batch_live = false;
  ctx->recs.clear();
  ctx->ev_used = 0;
  uint32_t* hS = (uint32_t*)ctx->hSmall.p;
  uint32_t* dS = (uint32_t*)ctx->small.p;  // [0]=bad offsets, [1]=retry count, [4..7]=CSum total, [8..9]=stmt totals

  // ---- input ----
  uint64_t off_first = 0, off_last = 0;
  const bool off32_in = (flags & NUTDB_F_OFFSETS32) != 0;  // the offsets are 32-bit words (half the upload)
  const size_t off_size = off32_in ? 4 : 8;
  const uint8_t* const off_bytes = reinterpret_cast<const uint8_t*>(stmt_off);
  if (dev_in) {
    hS[0] = hS[1] = hS[2] = hS[3] = 0;
    CK(cudaMemcpyAsync(hS, off_bytes, off_size, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hS + 2, off_bytes + off_size * n_stmt, off_size, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::memcpy(&off_first, hS, off_size);
    std::memcpy(&off_last, hS + 2, off_size);
  } else if (off32_in) {
    off_first = reinterpret_cast<const uint32_t*>(stmt_off)[0];
    off_last = reinterpret_cast<const uint32_t*>(stmt_off)[n_stmt];
  } else {
    off_first = stmt_off[0];
    off_last = stmt_off[n_stmt];
  }
  if (off_last < off_first || off_last - off_first >= 0x7FFFFFFFull) {
    ctx->err = "statement offsets must ascend and span fewer than 2^31 bytes";
    return NUTDB_E_ARG;
  }
  const uint32_t n = (uint32_t)(off_last - off_first);
  const uint32_t ntiles = (n + LEX_TILE - 1) / LEX_TILE;
  const size_t nchunks = (size_t)ntiles * LEX_THREADS;

  CK(cudaEventRecord(ctx->ev[0], st));
  ENSURE_DEV(off32, 4 * ((size_t)nstmt + 1));
  ENSURE_DEV(bitmap, 4 * (nchunks + 520));  // (k_lex4's tiles start at a range, not at a multiple of 8 KB: one tile of slack)
  ENSURE_DEV(stmt, sizeof(NutdbStmt) * ((size_t)nstmt + 1));
  ENSURE_DEV(stmtTokBegin, 4 * ((size_t)nstmt + 1));
  ENSURE_DEV(stmtTokEnd, 4 * ((size_t)nstmt + 1));
  const uint8_t* dText = nullptr;
  const uint64_t* dOff = nullptr;
  if (dev_in) {
    dOff = stmt_off;
    const uint8_t* p = sql + off_first;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      dText = p;
    } else {  // misaligned device input: one device-to-device copy into our aligned buffer
      ENSURE_DEV(text, (size_t)ntiles * LEX_TILE + 16);
      CK(cudaMemcpyAsync(ctx->text.p, p, n, cudaMemcpyDeviceToDevice, st));
      dText = (const uint8_t*)ctx->text.p;
    }
  } else {
    ENSURE_DEV(text, (size_t)ntiles * LEX_TILE + 16);
    ENSURE_DEV(off64, 8 * ((size_t)nstmt + 1));
    if (n) CK(cudaMemcpyAsync(ctx->text.p, sql + off_first, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->off64.p, stmt_off, off_size * ((size_t)nstmt + 1), cudaMemcpyHostToDevice, st));
    dText = (const uint8_t*)ctx->text.p;
    dOff = (const uint64_t*)ctx->off64.p;
  }
  CK(cudaEventRecord(ctx->ev[1], st));

  // ---- lexer ----
  uint32_t tok_cap = 0;
  int attempt = 0;
  bool use_lookback = ctx->force_lookback;  // k_lex3 instead of k_lex4 (a statement too long to cut the batch around)
  // bytes per range of k_lex4: at most L4_MAX_RANGES ranges, at least one tile each
  const uint32_t range_target = std::max<uint32_t>(L3_TILE, (n + L4_MAX_RANGES - 1) / L4_MAX_RANGES + 127u) & ~127u;
  const uint32_t n_readable = n;  // bytes of dText that may be read: bulk copies move whole 16-byte pieces below this
run_again:
  ENSURE_DEV(puntFlag, 4 * ((size_t)nstmt + 1));
  CK(cudaMemsetAsync(ctx->puntFlag.p, 0, 4 * ((size_t)nstmt + 1), st));
  ENSURE_DEV(firstStmt, 4 * (nchunks + 8));
  CK(cudaMemsetAsync(ctx->firstStmt.p, 0xFF, 4 * (nchunks + 8), st));
  CK(cudaMemsetAsync(ctx->bitmap.p, 0, 4 * (nchunks + 520), st));
  CK(cudaMemsetAsync(dS, 0, 128, st));
  {
    const uint32_t blocks = (uint32_t)(((uint64_t)nstmt + 1 + 255) / 256);
    if (off32_in)
      LAUNCH("k_prep", k_prep<uint32_t><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(dOff), nstmt, (uint64_t)n,
                                                              (uint32_t*)ctx->off32.p, (uint32_t*)ctx->bitmap.p,
                                                              (uint32_t*)ctx->firstStmt.p, dS));
    else
      LAUNCH("k_prep", k_prep<uint64_t><<<blocks, 256, 0, st>>>(dOff, nstmt, (uint64_t)n, (uint32_t*)ctx->off32.p, (uint32_t*)ctx->bitmap.p,
                                                              (uint32_t*)ctx->firstStmt.p, dS));
  }
  uint32_t ntok = 0;
  ctx->n_punt = 0;
  if (n > 0 && lex_only) {  // verify mode: the thread-per-chunk walker (exact for every input, emits every token)
    ENSURE_DEV(localA, 4 * nchunks);
    ENSURE_DEV(localB, 4 * nchunks);
    ENSURE_DEV(localC, 16 * nchunks);
    ENSURE_DEV(tileA, 4 * (size_t)ntiles);
    ENSURE_DEV(tileB, 4 * (size_t)ntiles);
    ENSURE_DEV(tileC, 16 * (size_t)ntiles);
    ENSURE_DEV(tilePrefC, 16 * (size_t)ntiles);
    ENSURE_DEV(entA, ntiles);
    ENSURE_DEV(entB, ntiles);
    const uint32_t* bm = (const uint32_t*)ctx->bitmap.p;
    LAUNCH("k_lex_A", k_lex_A<<<ntiles, LEX_THREADS, 0, st>>>(dText, bm, n, ctx->dLex, (uint32_t*)ctx->localA.p, (uint32_t*)ctx->tileA.p));
    LAUNCH("k_scan_A", k_scan_vec8<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)ctx->tileA.p, (uint8_t*)ctx->entA.p, ntiles));
    LAUNCH("k_lex_B", k_lex_B<<<ntiles, LEX_THREADS, 0, st>>>(dText, bm, n, ctx->dLex, (const uint32_t*)ctx->localA.p,
                                            (const uint8_t*)ctx->entA.p, (uint32_t*)ctx->localB.p,
                                            (uint32_t*)ctx->tileB.p));
    LAUNCH("k_scan_B", k_scan_vec8<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)ctx->tileB.p, (uint8_t*)ctx->entB.p, ntiles));
    if (lex_only)
      LAUNCH("k_lex_C", k_lex_C<true><<<ntiles, LEX_THREADS, 0, st>>>(
                            dText, bm, n, ctx->dLex, (const uint32_t*)ctx->localA.p, (const uint8_t*)ctx->entA.p,
                            (const uint32_t*)ctx->localB.p, (const uint8_t*)ctx->entB.p, (uint4*)ctx->localC.p,
                            (uint4*)ctx->tileC.p));
    else
      LAUNCH("k_lex_C", k_lex_C<false><<<ntiles, LEX_THREADS, 0, st>>>(
                            dText, bm, n, ctx->dLex, (const uint32_t*)ctx->localA.p, (const uint8_t*)ctx->entA.p,
                            (const uint32_t*)ctx->localB.p, (const uint8_t*)ctx->entB.p, (uint4*)ctx->localC.p,
                            (uint4*)ctx->tileC.p));
    LAUNCH("k_scan_C", k_scan_tiles<CSumOp><<<1, SCAN_THREADS, 0, st>>>((const uint4*)ctx->tileC.p, (uint4*)ctx->tilePrefC.p, ntiles,
                                                     (uint4*)(dS + 4)));
    CK(cudaMemcpyAsync(hS, dS, 64, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hS[0]) {
      ctx->err = "statement offsets must ascend";
      return NUTDB_E_ARG;
    }
    ntok = hS[4];
    ENSURE_DEV(tokType, (size_t)ntok + 16);
    ENSURE_DEV(tokKw, (size_t)ntok + 16);
    ENSURE_DEV(tokStart, 4 * ((size_t)ntok + 4));
    ENSURE_DEV(tokEnd, 4 * ((size_t)ntok + 4));
    DevSink sink{(uint8_t*)ctx->tokType.p, (uint32_t*)ctx->tokStart.p, (uint32_t*)ctx->tokEnd.p, (uint8_t*)ctx->tokKw.p,
                 ntok, (uint32_t*)ctx->stmtTokBegin.p, (uint32_t*)ctx->stmtTokEnd.p, (const uint32_t*)ctx->off32.p, nstmt};
    if (lex_only)
      LAUNCH("k_lex_D", k_lex_D<true><<<ntiles, LEX_THREADS, 0, st>>>(
                            dText, bm, n, ctx->dLex, (const uint32_t*)ctx->localA.p, (const uint8_t*)ctx->entA.p,
                            (const uint32_t*)ctx->localB.p, (const uint8_t*)ctx->entB.p, (const uint4*)ctx->localC.p,
                            (const uint4*)ctx->tilePrefC.p, sink));
    else
      LAUNCH("k_lex_D", k_lex_D<false><<<ntiles, LEX_THREADS, 0, st>>>(
                            dText, bm, n, ctx->dLex, (const uint32_t*)ctx->localA.p, (const uint8_t*)ctx->entA.p,
                            (const uint32_t*)ctx->localB.p, (const uint8_t*)ctx->entB.p, (const uint4*)ctx->localC.p,
                            (const uint4*)ctx->tilePrefC.p, sink));
  } else if (n > 0) {
    // ---- single-pass lexer (lex3_core.cuh) + exact walker for the statements it flags ----
    // Everything up to the dense node layout is issued without the host looking at a device-side count: token arrays
    // are sized by an estimate (grow-only; an overflow is detected at the one synchronisation and the batch re-run),
    // list lengths are read by the kernels themselves.
    // k_lex4 lexes ranges of whole statements independently; k_lex3 (one token stream, look-back scans between tiles)
    // takes over when a statement is too long to cut the batch around it.
    const uint32_t ntiles3 = (n + L3_TILE - 1) / L3_TILE;
    const size_t nwin = (size_t)ntiles3 * L3_WIN + L4_MAX_RANGES + 8;  // (+ one slot per range, see k_lex4)
    ENSURE_DEV(winIdx, 4 * (nwin + 8));
    ENSURE_DEV(winHas, 4 * (nwin + 8));
    ENSURE_DEV(winEof, 4 * (nwin + 8));
    ENSURE_DEV(rangeByte, 4 * (L4_MAX_RANGES + 2));
    ENSURE_DEV(rangeStmt, 4 * (L4_MAX_RANGES + 2));
    ENSURE_DEV(rangeTokBase, 4 * (L4_MAX_RANGES + 2));
    ENSURE_DEV(rangeCount, 8 * (L4_MAX_RANGES + 2));
    ENSURE_DEV(rangeDense, 8 * (L4_MAX_RANGES + 2));
    ENSURE_DEV(puntList, 4 * ((size_t)nstmt + 1));
    ENSURE_DEV(puntCounts, 8 * ((size_t)nstmt + 1));
    ENSURE_DEV(puntOffs, 8 * ((size_t)nstmt + 1));
    const uint32_t nb = (nstmt + PUNT_THREADS * PUNT_PER_THREAD - 1) / (PUNT_THREADS * PUNT_PER_THREAD);
    ENSURE_DEV(puntBlockCount, 8 * ((size_t)nb + 1));
    ENSURE_DEV(puntBlockPref, 8 * ((size_t)nb + 1));
    // ~0.2 tokens per byte on query logs; dense text ("((((") overflows the estimate once and grows the buffers
    const uint32_t cap_num = ctx->tok_cap_num;  // sixteenths of a token per byte
    {
      size_t want = (size_t)n * cap_num / 16 + 2 * (size_t)nstmt + 64 * (size_t)L4_MAX_RANGES + 4096;
      if (want > (size_t)n + nstmt + 64 * (size_t)L4_MAX_RANGES + 4096) want = (size_t)n + nstmt + 64 * (size_t)L4_MAX_RANGES + 4096;
      if (want < ctx->tok_cap_min) want = ctx->tok_cap_min;
      if (want > 0xFFFFFFE0ull) want = 0xFFFFFFE0ull;
      ENSURE_DEV(tokType, want + 16);
      ENSURE_DEV(tokKw, want + 16);
      ENSURE_DEV(tokStart, 4 * (want + 4));
      ENSURE_DEV(tokEnd, 4 * (want + 4));
      tok_cap = (uint32_t)std::min<size_t>({(size_t)0xFFFFFFE0u, ctx->tokType.cap - 16, ctx->tokKw.cap - 16,
                                            ctx->tokStart.cap / 4 - 4, ctx->tokEnd.cap / 4 - 4});
    }
    const uint32_t* bm = (const uint32_t*)ctx->bitmap.p;
    // counters: [0] flagged statements, [1] bound on their tokens, [2] start of the exact lexer's token region (k_lex3: the
    // number of main-region tokens), [3] tile / range ticket, [4] a segment overflowed
    uint32_t* counters = dS + 14;
    uint32_t* cuts_info = dS + 20;  // {ranges, end of the last segment, longest range in bytes}
    Lex3Out lo3{(uint8_t*)ctx->tokType.p, (uint32_t*)ctx->tokStart.p, (uint32_t*)ctx->tokEnd.p, (uint8_t*)ctx->tokKw.p, tok_cap,
                (uint32_t*)ctx->winIdx.p, (uint32_t*)ctx->winHas.p, (uint32_t*)ctx->winEof.p, (const uint32_t*)ctx->off32.p, nstmt,
                (uint32_t*)ctx->puntFlag.p, counters, (const uint32_t*)ctx->firstStmt.p, nullptr, nullptr};
    const uint32_t* extra_base_ptr;
    if (!use_lookback) {
      LAUNCH("k_cuts", k_cuts<<<1, 1024, 0, st>>>((const uint32_t*)ctx->off32.p, nstmt, n, range_target, cap_num,
                                                  (uint32_t*)ctx->rangeByte.p, (uint32_t*)ctx->rangeStmt.p,
                                                  (uint32_t*)ctx->rangeTokBase.p, cuts_info, dS));
      Lex4Ranges rg{(const uint32_t*)ctx->rangeByte.p, (const uint32_t*)ctx->rangeTokBase.p, (uint2*)ctx->rangeCount.p, cuts_info};
      const uint32_t grid = std::min<uint32_t>(std::min<uint32_t>(L4_MAX_RANGES, (n + range_target - 1) / range_target),
                                               (uint32_t)ctx->sm_count * L4_MINBLOCKS);
      LAUNCH("k_lex4", k_lex4<<<grid, L4_THREADS, sizeof(Lex4Shared), st>>>(dText, bm, n, n_readable, ctx->dLex, rg, lo3, dS));
      // (dense positions of the segments, for callers that want the token arrays)
      LAUNCH("k_scan_R", k_scan_tiles<U2AddOp><<<1, SCAN_THREADS, 0, st>>>((const uint2*)ctx->rangeCount.p, (uint2*)ctx->rangeDense.p,
                                                                           L4_MAX_RANGES, (uint2*)(dS + 24), cuts_info));
      extra_base_ptr = cuts_info + 1;
    } else {
      for (DevBuf* d : {&ctx->descFn, &ctx->descA, &ctx->descI, &ctx->descB, &ctx->descC}) {  // look-back descriptors: tags start at epoch 0
        const int rc = ensure_dev_zeroed(ctx, *d, 8 * (size_t)ntiles3, st);
        if (rc != NUTDB_OK) return rc;
      }
      Lex3Desc desc{(unsigned long long*)ctx->descFn.p, (unsigned long long*)ctx->descA.p, (unsigned long long*)ctx->descI.p,
                    (unsigned long long*)ctx->descB.p, (unsigned long long*)ctx->descC.p};
      if (ctx->debug_tiles) {
        ENSURE_DEV(dbgTiles, 16 * (size_t)ntiles3 + 16);
        lo3.dbg = (uint32_t*)ctx->dbgTiles.p;
        ctx->dbg_ntiles = ntiles3;
      }
      if (ctx->debug_timing) {
        ENSURE_DEV(dbgTim, 24 * 8 * (size_t)ntiles3 + 16);
        CK(cudaMemsetAsync(ctx->dbgTim.p, 0, 24 * 8 * (size_t)ntiles3, st));
        lo3.tim = (unsigned long long*)ctx->dbgTim.p;
        ctx->dbg_ntiles = ntiles3;
      }
      ctx->epoch = (ctx->epoch + 1u) & 0x07FFFFFFu;  // (the tags hold 27 bits of it)
      if (ctx->epoch == 0u) ctx->epoch = 1u;
      // one range: the whole batch, window slots unshifted, tokens dense from 0
      const uint32_t one[8] = {0u, n, 0u, nstmt, 0u, tok_cap, 1u, 0u};
      std::memcpy(hS + 40, one, sizeof(one));
      CK(cudaMemcpyAsync(ctx->rangeByte.p, hS + 40, 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->rangeStmt.p, hS + 42, 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->rangeTokBase.p, hS + 44, 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(cuts_info, hS + 46, 4, cudaMemcpyHostToDevice, st));
#ifndef L3_GRID_PER_SM
#define L3_GRID_PER_SM L3_MINBLOCKS
#endif
      // (+ block 0, the scanner of the token counts)
      const uint32_t grid = 1u + std::min<uint32_t>(ntiles3, (uint32_t)ctx->sm_count * L3_GRID_PER_SM);
      LAUNCH("k_lex3", k_lex3<<<grid, L3_THREADS, sizeof(Lex3Shared), st>>>(dText, bm, n, n_readable, ntiles3, ctx->dLex, desc,
                                                                            ctx->epoch, lo3, dS));
      extra_base_ptr = counters + 2;
    }
    // the flagged statements in ascending order (deterministic layout of the extra token region), then the exact walker
    LAUNCH("k_punt_count", k_punt_list<false><<<nb, PUNT_THREADS, 0, st>>>((const uint32_t*)ctx->puntFlag.p, nstmt,
                                                                           (uint2*)ctx->puntBlockCount.p, nullptr, nullptr, counters));
    LAUNCH("k_scan_PB", k_scan_tiles<U2AddOp><<<1, SCAN_THREADS, 0, st>>>((const uint2*)ctx->puntBlockCount.p,
                                                                          (uint2*)ctx->puntBlockPref.p, nb, nullptr));
    LAUNCH("k_punt_scatter", k_punt_list<true><<<nb, PUNT_THREADS, 0, st>>>((const uint32_t*)ctx->puntFlag.p, nstmt, nullptr,
                                                                            (const uint2*)ctx->puntBlockPref.p,
                                                                            (uint32_t*)ctx->puntList.p, counters));
    const uint32_t xgrid = std::min<uint32_t>((nstmt + 127) / 128, 4096u);
    ExactSink xs{nullptr, nullptr, nullptr, nullptr, 0};
    LAUNCH("k_lex_exact_count", k_lex_exact<false><<<xgrid, 128, 0, st>>>(
                                    dText, (const uint32_t*)ctx->off32.p, ctx->dLex, (const uint32_t*)ctx->puntList.p, counters,
                                    (uint2*)ctx->puntCounts.p, nullptr, nullptr, xs, nullptr, nullptr, nullptr));
    LAUNCH("k_scan_P", k_scan_tiles<U2AddOp><<<1, SCAN_THREADS, 0, st>>>((const uint2*)ctx->puntCounts.p, (uint2*)ctx->puntOffs.p,
                                                                         nstmt, (uint2*)(dS + 10), counters));
    xs = ExactSink{lo3.type, lo3.start, lo3.end, lo3.kw, tok_cap};
    LAUNCH("k_lex_exact_emit", k_lex_exact<true><<<xgrid, 128, 0, st>>>(
                                   dText, (const uint32_t*)ctx->off32.p, ctx->dLex, (const uint32_t*)ctx->puntList.p, counters, nullptr,
                                   (const uint2*)ctx->puntOffs.p, extra_base_ptr, xs, (uint32_t*)ctx->stmtTokBegin.p,
                                   (uint32_t*)ctx->stmtTokEnd.p, (uint32_t*)ctx->puntFlag.p));
  }
  CK(cudaEventRecord(ctx->ev[2], st));

  // ---- parser ----
  uint64_t n_node = 0, n_err = 0;
  uint32_t ext_cap = 0;
  // 8-byte statement records on the wire: only where nobody asked for tokens (the record's token fields are dropped)
  const bool wire_stmt = (flags & NUTDB_F_WIRE_STMT) && (flags & NUTDB_F_NO_TOKENS) && !lex_only && nstmt > 0;
  const bool native_lex = n > 0 && !lex_only;
  if (nstmt > 0) {
    // compact-node scratch: a disjoint range of tok_count + NODE_SLACK slots per statement (see node_slot)
    const size_t scratch_nodes = (native_lex ? (size_t)tok_cap : (size_t)ntok) + (size_t)NODE_SLACK * 2 * ((size_t)nstmt + 1) + 4;
    if (!lex_only) ENSURE_DEV(scratch, 8 * scratch_nodes);
    ENSURE_DEV(retryList, 8 * ((size_t)nstmt + 1));
    ENSURE_DEV(slowList, 4 * ((size_t)nstmt + 1));
    ENSURE_DEV(slowList2, 4 * ((size_t)nstmt + 1));
    StmtToks stoks{native_lex ? (const uint32_t*)ctx->winIdx.p : nullptr, (const uint32_t*)ctx->winHas.p,
                   (const uint32_t*)ctx->winEof.p, (const uint32_t*)ctx->puntFlag.p, (const uint32_t*)ctx->stmtTokBegin.p,
                   (const uint32_t*)ctx->stmtTokEnd.p, (const uint32_t*)ctx->rangeStmt.p, dS + 20,
                   native_lex ? tok_cap : 0xFFFFFFFFu};
    LAUNCH("k_parse_fast", k_parse_fast<<<(nstmt + FAST_THREADS - 1) / FAST_THREADS, FAST_THREADS, FAST_DYN_SMEM, st>>>(
                               dText, (const uint32_t*)ctx->off32.p, nstmt, ntok, (const uint8_t*)ctx->tokType.p,
                               (const uint32_t*)ctx->tokStart.p, (const uint32_t*)ctx->tokEnd.p,
                               (const uint8_t*)ctx->tokKw.p, stoks, (NutdbStmt*)ctx->stmt.p, (uint2*)ctx->scratch.p,
                               (uint32_t*)ctx->slowList.p, dS + 2, (const uint32_t*)ctx->puntFlag.p, lex_only ? 1 : 0,
                               native_lex ? tok_cap : (uint32_t)min((size_t)0xFFFFFFF0u, (size_t)ntok + 16), ctx->dFast,
                               native_lex ? (use_lookback ? dS + 16 : dS + 21) : nullptr, native_lex ? dS + 10 : nullptr, dS));
    if (!lex_only) {
      const uint32_t gcap = nstmt + 1u;  // (a group holds at least one statement)
      ENSURE_DEV(wideOrder, 4 * ((size_t)nstmt + WIDE_POOL));
      ENSURE_DEV(wideGroups, 8 * (size_t)gcap);
      ENSURE_DEV(widePools, 8 * ((size_t)nstmt / WIDE_POOL + 2));
      LAUNCH("k_wide_order", k_wide_order<<<(uint32_t)min(((size_t)nstmt + WIDE_POOL - 1) / WIDE_POOL, (size_t)ctx->sm_count * 8), WIDE_POOL, 0, st>>>(
                                 (const uint32_t*)ctx->slowList.p, dS + 2, (const uint8_t*)ctx->tokType.p,
                                 (const NutdbStmt*)ctx->stmt.p, (uint32_t*)ctx->wideOrder.p, (uint2*)ctx->wideGroups.p, gcap, dS + 30,
                                 (uint2*)ctx->widePools.p, dS + 3));
      const uint32_t wgrid = (uint32_t)min(((size_t)nstmt + WIDE_THREADS / 32 - 1) / (WIDE_THREADS / 32), (size_t)ctx->sm_count * WIDE_CTAS_PER_SM);
      LAUNCH("k_parse_wide", k_parse_wide<<<wgrid, WIDE_THREADS, 0, st>>>(
                                 dText, (const uint32_t*)ctx->off32.p, (const uint32_t*)ctx->wideOrder.p,
                                 (const uint2*)ctx->wideGroups.p, gcap, dS + 30, (const uint2*)ctx->widePools.p, dS + 3,
                                 (const uint8_t*)ctx->tokType.p, (const uint32_t*)ctx->tokStart.p, (const uint32_t*)ctx->tokEnd.p,
                                 (const uint8_t*)ctx->tokKw.p, ctx->dFast, (NutdbStmt*)ctx->stmt.p, (uint2*)ctx->scratch.p,
                                 (uint32_t*)ctx->slowList2.p, dS + 27, nstmt, (const uint32_t*)ctx->puntFlag.p, dS + 29, dS + 23));
    }
    if (!lex_only)
      LAUNCH("k_parse", k_parse<<<(nstmt + PARSE_THREADS - 1) / PARSE_THREADS, PARSE_THREADS, 0, st>>>(
                            dText, (const uint32_t*)ctx->off32.p, (const uint32_t*)ctx->slowList2.p, 0u,
                            (const uint8_t*)ctx->tokType.p, (const uint32_t*)ctx->tokStart.p,
                            (const uint32_t*)ctx->tokEnd.p, (const uint8_t*)ctx->tokKw.p, ctx->dPar,
                            (NutdbStmt*)ctx->stmt.p, (uint2*)ctx->scratch.p, (uint2*)ctx->retryList.p, dS + 1, nstmt,
                            (const uint32_t*)ctx->puntFlag.p, dS + 27));
    const uint32_t stiles = (nstmt + FIN_THREADS - 1) / FIN_THREADS;
    ENSURE_DEV(tileS, 8 * (size_t)stiles);
    ENSURE_DEV(tilePrefS, 8 * (size_t)stiles);
    LAUNCH("k_stmt_sums", k_stmt_sums<<<stiles, FIN_THREADS, 0, st>>>((const NutdbStmt*)ctx->stmt.p, nstmt, (uint2*)ctx->tileS.p));
    LAUNCH("k_scan_S", k_scan_tiles<U2AddOp><<<1, SCAN_THREADS, 0, st>>>((const uint2*)ctx->tileS.p, (uint2*)ctx->tilePrefS.p, stiles,
                                                      (uint2*)(dS + 8)));
    // ---- the one synchronisation of the call: counts the host needs to size what follows ----
    CK(cudaMemcpyAsync(hS, dS, 128, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hS[0] & 1u) {
      ctx->err = "statement offsets must ascend and lie inside the batch";
      return NUTDB_E_ARG;
    }
    if (native_lex) {
      // k_lex4: hS[21] = end of the segments (where the exact lexer's region begins), hS[24] = tokens in them;
      // k_lex3: hS[16] = tokens of the main region = start of the exact lexer's region
      const uint64_t seg_end = use_lookback ? hS[16] : hS[21], n_main = use_lookback ? hS[16] : hS[24], n_extra = hS[10];
      ctx->n_punt = hS[14];
      if (seg_end + n_extra >= 0xFFFFFFE0ull) {
        ctx->err = "too many tokens in one batch";
        return NUTDB_E_ARG;
      }
      const bool seg_overflow = !use_lookback && hS[18] != 0;
      const bool giant = !use_lookback && (uint64_t)hS[22] > std::max<uint64_t>(4u << 20, 8ull * range_target);
      if (ctx->debug_sync) fprintf(stderr, "nutdb_gpu: attempt %d n=%u nstmt=%u tok_cap=%u seg_end=%llu n_main=%llu n_extra=%llu overflow=%u longest=%u cap_num=%u lookback=%d\n", attempt, n, nstmt, tok_cap, (unsigned long long)seg_end, (unsigned long long)n_main, (unsigned long long)n_extra, hS[18], hS[22], ctx->tok_cap_num, (int)use_lookback);
      if (seg_overflow || giant || seg_end + n_extra > tok_cap) {  // an estimate was too small / the cuts do not fit: run again
        if (attempt >= 4) {
          ctx->err = "token buffers kept overflowing";
          return NUTDB_E_NOMEM;
        }
        if (giant) use_lookback = true;  // a statement too long to cut around: one token stream, look-back scans
        else if (seg_overflow) ctx->tok_cap_num = std::min<uint32_t>(16u, ctx->tok_cap_num * 2u);
        if (seg_end + n_extra > tok_cap) ctx->tok_cap_min = (size_t)(seg_end + n_extra) + (size_t)(seg_end + n_extra) / 16 + 1024;
        attempt++;
        goto run_again;
      }
      ntok = (uint32_t)(n_main + n_extra);
      ctx->last_lookback = use_lookback;
    }
    ctx->n_slow = hS[27];           // statements the exact automaton parsed
    ctx->n_wide = hS[2] - hS[27];   // statements the wide table-driven pass parsed
    const uint32_t nretry = hS[1];
    if (nretry > 0) {
      // deep statements: per-statement stack and node ranges sized from their token counts
      ENSURE_HOST(hRetry, 8 * (size_t)nretry);
      const uint2* hl = (const uint2*)ctx->hRetry.p;
      CK(cudaMemcpyAsync(ctx->hRetry.p, ctx->retryList.p, 8 * (size_t)nretry, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      std::vector<uint64_t> noff(nretry + 1), soff(nretry + 1);
      noff[0] = soff[0] = 0;
      for (uint32_t i = 0; i < nretry; i++) {
        uint64_t tc = hl[i].y;
        noff[i + 1] = noff[i] + 2 * tc + 8;
        soff[i + 1] = soff[i] + 8 * tc + 64;
      }
      if (noff[nretry] >= 0xFFFFFFF0ull) {
        ctx->err = "deep-statement scratch exceeds 2^32 nodes";
        return NUTDB_E_NOMEM;
      }
      ENSURE_DEV(retryNodeOff, 8 * ((size_t)nretry + 1));
      ENSURE_DEV(retryStackOff, 8 * ((size_t)nretry + 1));
      ENSURE_DEV(retryNodes, 8 * (size_t)noff[nretry] + 64);
      ENSURE_DEV(retryStack, 4 * (size_t)soff[nretry] + 64);
      CK(cudaMemcpyAsync(ctx->retryNodeOff.p, noff.data(), 8 * ((size_t)nretry + 1), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->retryStackOff.p, soff.data(), 8 * ((size_t)nretry + 1), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));  // noff/soff are stack vectors
      uint32_t rspread = 1;
      while (rspread < 32u && (uint64_t)nretry * rspread * 2u <= (uint64_t)ctx->sm_count * 1024u) rspread <<= 1;
      LAUNCH("k_parse_retry", k_parse_retry<<<(uint32_t)(((uint64_t)nretry * rspread + PARSE_THREADS - 1) / PARSE_THREADS), PARSE_THREADS, 0, st>>>(
          dText, (const uint32_t*)ctx->off32.p, (const uint8_t*)ctx->tokType.p, (const uint32_t*)ctx->tokStart.p,
          (const uint32_t*)ctx->tokEnd.p, (const uint8_t*)ctx->tokKw.p, ctx->dPar, (NutdbStmt*)ctx->stmt.p,
          (const uint2*)ctx->retryList.p, nretry, (const uint64_t*)ctx->retryNodeOff.p,
          (const uint64_t*)ctx->retryStackOff.p, (uint2*)ctx->retryNodes.p, (uint32_t*)ctx->retryStack.p, rspread));
      LAUNCH("k_stmt_sums", k_stmt_sums<<<stiles, FIN_THREADS, 0, st>>>((const NutdbStmt*)ctx->stmt.p, nstmt, (uint2*)ctx->tileS.p));
      LAUNCH("k_scan_S", k_scan_tiles<U2AddOp><<<1, SCAN_THREADS, 0, st>>>((const uint2*)ctx->tileS.p, (uint2*)ctx->tilePrefS.p,
                                                                           stiles, (uint2*)(dS + 8)));
      CK(cudaMemcpyAsync(hS, dS, 64, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    n_node = hS[8];
    n_err = hS[9];
    ENSURE_DEV(nodes, 4 * (n_node + 4));
    ENSURE_DEV(errs, 32 * (n_err + 1));
    if (wire_stmt) ENSURE_DEV(wstmt, 8 * ((size_t)nstmt + 1));
    ext_cap = std::max<uint32_t>(n / 64u, 1024u);  // side table of the nodes that do not fit the 32-bit wire form
    ENSURE_DEV(extNodes, 16 * (size_t)ext_cap);
    LAUNCH("k_finalize", k_finalize<<<stiles, FIN_THREADS, 0, st>>>((NutdbStmt*)ctx->stmt.p, nstmt, (const uint2*)ctx->tilePrefS.p,
                                               (const uint2*)ctx->scratch.p, (const uint2*)ctx->retryNodes.p,
                                               (const uint32_t*)ctx->tokStart.p, (const uint32_t*)ctx->tokEnd.p,
                                               (const uint32_t*)ctx->puntFlag.p, (uint32_t*)ctx->nodes.p,
                                               (uint4*)ctx->errs.p, dS + 26, (uint4*)ctx->extNodes.p, ext_cap,
                                               wire_stmt ? (unsigned long long*)ctx->wstmt.p : nullptr, dS + 28));
  } else {
    CK(cudaMemcpyAsync(hS, dS, 128, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  // k_lex4 left gaps between the ranges' token segments: callers that see the token arrays get dense ones
  const uint8_t* dTokType = (const uint8_t*)ctx->tokType.p;
  const uint8_t* dTokKw = (const uint8_t*)ctx->tokKw.p;
  const uint32_t* dTokStart = (const uint32_t*)ctx->tokStart.p;
  const uint32_t* dTokEnd = (const uint32_t*)ctx->tokEnd.p;
  if (native_lex && !use_lookback && nstmt > 0 && !(flags & NUTDB_F_NO_TOKENS)) {
    ENSURE_DEV(tokTypeD, (size_t)ntok + 16);
    ENSURE_DEV(tokKwD, (size_t)ntok + 16);
    ENSURE_DEV(tokStartD, 4 * ((size_t)ntok + 4));
    ENSURE_DEV(tokEndD, 4 * ((size_t)ntok + 4));
    const uint32_t slices = 8;
    LAUNCH("k_tok_compact", k_tok_compact<<<(L4_MAX_RANGES + 1) * slices, 256, 0, st>>>(
                                dTokType, dTokStart, dTokEnd, dTokKw, (uint8_t*)ctx->tokTypeD.p, (uint32_t*)ctx->tokStartD.p,
                                (uint32_t*)ctx->tokEndD.p, (uint8_t*)ctx->tokKwD.p, (const uint32_t*)ctx->rangeTokBase.p,
                                (const uint2*)ctx->rangeCount.p, (const uint2*)ctx->rangeDense.p, dS + 20, dS + 10, slices));
    LAUNCH("k_stmt_tok_remap", k_stmt_tok_remap<<<(nstmt + 255) / 256, 256, 0, st>>>(
                                   (NutdbStmt*)ctx->stmt.p, nstmt, (const uint32_t*)ctx->rangeStmt.p, (const uint32_t*)ctx->rangeTokBase.p,
                                   (const uint2*)ctx->rangeCount.p, (const uint2*)ctx->rangeDense.p, dS + 20));
    dTokType = (const uint8_t*)ctx->tokTypeD.p;
    dTokKw = (const uint8_t*)ctx->tokKwD.p;
    dTokStart = (const uint32_t*)ctx->tokStartD.p;
    dTokEnd = (const uint32_t*)ctx->tokEndD.p;
  }
  CK(cudaEventRecord(ctx->ev[3], st));

  // ---- output ----
  out->n_stmt = nstmt;
  out->n_tok = ntok;
  out->n_node = n_node;
  out->n_err = n_err;
  hS[48] = 0;
  hS[49] = 0;
  if (n_node) CK(cudaMemcpyAsync(hS + 48, dS + 26, 4, cudaMemcpyDeviceToHost, st));  // side-table entries (almost always none)
  if (wire_stmt) CK(cudaMemcpyAsync(hS + 49, dS + 28, 4, cudaMemcpyDeviceToHost, st));  // a count did not fit the wire record
  if (!(flags & NUTDB_F_NO_HOST_COPY)) {
    ENSURE_HOST(hStmt, sizeof(NutdbStmt) * ((size_t)nstmt + 1));
    ENSURE_HOST(hNode, 4 * (n_node + 4));
    ENSURE_HOST(hErr, 32 * (n_err + 1));
    if (nstmt && wire_stmt) CK(cudaMemcpyAsync(ctx->hStmt.p, ctx->wstmt.p, 8 * (size_t)nstmt, cudaMemcpyDeviceToHost, st));
    else if (nstmt) CK(cudaMemcpyAsync(ctx->hStmt.p, ctx->stmt.p, sizeof(NutdbStmt) * (size_t)nstmt, cudaMemcpyDeviceToHost, st));
    if (n_node) CK(cudaMemcpyAsync(ctx->hNode.p, ctx->nodes.p, 4 * n_node, cudaMemcpyDeviceToHost, st));
    if (n_err) CK(cudaMemcpyAsync(ctx->hErr.p, ctx->errs.p, 32 * n_err, cudaMemcpyDeviceToHost, st));
    if (wire_stmt) out->wstmt = (const uint64_t*)ctx->hStmt.p;
    else out->stmt = (const NutdbStmt*)ctx->hStmt.p;
    out->pnode = (const uint32_t*)ctx->hNode.p;
    out->err = (const NutdbError*)ctx->hErr.p;
    if (!(flags & NUTDB_F_NO_TOKENS)) {
      ENSURE_HOST(hTokType, (size_t)ntok + 16);
      ENSURE_HOST(hTokKw, (size_t)ntok + 16);
      ENSURE_HOST(hTokStart, 4 * ((size_t)ntok + 4));
      ENSURE_HOST(hTokEnd, 4 * ((size_t)ntok + 4));
      if (ntok) {
        CK(cudaMemcpyAsync(ctx->hTokType.p, dTokType, ntok, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->hTokKw.p, dTokKw, ntok, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->hTokStart.p, dTokStart, 4 * (size_t)ntok, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->hTokEnd.p, dTokEnd, 4 * (size_t)ntok, cudaMemcpyDeviceToHost, st));
      }
      out->tok_type = (const uint8_t*)ctx->hTokType.p;
      out->tok_kw = (const uint8_t*)ctx->hTokKw.p;
      out->tok_start = (const uint32_t*)ctx->hTokStart.p;
      out->tok_end = (const uint32_t*)ctx->hTokEnd.p;
    }
  }
  CK(cudaEventRecord(ctx->ev[4], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  bool wire_ok = wire_stmt;
  if (wire_stmt && hS[49]) {  // (a statement with 2^30 nodes or tokens: the full records after all)
    wire_ok = false;
    if (!(flags & NUTDB_F_NO_HOST_COPY)) {
      CK(cudaMemcpy(ctx->hStmt.p, ctx->stmt.p, sizeof(NutdbStmt) * (size_t)nstmt, cudaMemcpyDeviceToHost));
      out->wstmt = nullptr;
      out->stmt = (const NutdbStmt*)ctx->hStmt.p;
    }
  }
  ctx->ext_count = 0;
  if (n_node && hS[48]) {  // nodes that did not fit the 32-bit wire form: their exact fields, sorted by node index
    if (hS[48] > ext_cap) {
      ctx->err = "node side table overflow";
      return NUTDB_E_NOMEM;
    }
    ctx->ext_count = hS[48];
    out->n_ext = hS[48];
    // (outputs left on the device: the table is fetched and sorted when somebody asks -- nutdb_gpu_batch_fetch_ext)
    if (!(flags & NUTDB_F_NO_HOST_COPY)) {
      const int rc = fetch_ext(ctx);
      if (rc != NUTDB_OK) return rc;
      out->ext = ctx->ext.data();
    }
  }
  cudaEventElapsedTime(&ctx->ms[0], ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->ms[1], ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->ms[2], ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&ctx->ms[3], ctx->ev[3], ctx->ev[4]);
  cudaEventElapsedTime(&ctx->ms[4], ctx->ev[0], ctx->ev[4]);
  ctx->kernel_ms.clear();
  for (const NutdbCtx::KRec& r : ctx->recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    ctx->kernel_ms.emplace_back(r.name, ms);
  }
  ctx->dev_view.stmt = ctx->stmt.p;
  ctx->dev_view.wstmt = wire_ok ? ctx->wstmt.p : nullptr;
  ctx->dev_view.tok_type = dTokType;
  ctx->dev_view.tok_start = dTokStart;
  ctx->dev_view.tok_end = dTokEnd;
  ctx->dev_view.tok_kw = dTokKw;
  ctx->dev_view.node = ctx->nodes.p;
  ctx->dev_view.err = ctx->errs.p;
  ctx->batch_live = true;
  out->impl = ctx;
  return NUTDB_OK;
}

void nutdb_gpu_batch_free(NutdbCtx* ctx, NutdbBatch* batch) {
  // Output buffers are owned by the context and re-used by the next batch (grow-only), so
  // releasing a batch only invalidates the caller's view.
  if (ctx) ctx->batch_live = false;
  if (batch) std::memset(batch, 0, sizeof(*batch));
}

}  // extern "C"

// ---- 64-bit checksum of a batch's device-resident outputs (nutdb_gpu_batch_hash) ----
__host__ __device__ __forceinline__ uint64_t hash_mix64(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash_term(uint32_t array, uint64_t index, uint32_t value) {
  return hash_mix64(hash_mix64(index + (uint64_t)(array + 1u) * 0x9E3779B97F4A7C15ull) ^ (uint64_t)value);
}
// array 0..: stmt words, tok_type bytes, tok_start, tok_end, tok_kw bytes, node words, err words.  One launch per array.
template <typename T>
__global__ void __launch_bounds__(256) k_hash(const T* __restrict__ a, uint64_t n, uint32_t array,
                                              unsigned long long* __restrict__ out) {
  uint64_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    acc += hash_term(array, i, (uint32_t)a[i]);
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if ((threadIdx.x & 31u) == 0 && acc != 0) atomicAdd(out, (unsigned long long)acc);
}

extern "C" {

int nutdb_gpu_batch_hash(const NutdbBatch* batch, uint64_t* out) {
  if (!batch || !out || !batch->impl) return NUTDB_E_ARG;
  NutdbCtx* ctx = (NutdbCtx*)batch->impl;
  if (!ctx->batch_live) return NUTDB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  {
    const int rc = ensure_dev(ctx, ctx->hashAcc, 8);
    if (rc != NUTDB_OK) return rc;
  }
  unsigned long long* acc = (unsigned long long*)ctx->hashAcc.p;
  CK(cudaMemsetAsync(acc, 0, 8, st));
  const int grid = 148 * 8;
  const NutdbBatchDevice& v = ctx->dev_view;
  if (batch->n_stmt) k_hash<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)v.stmt, batch->n_stmt * (sizeof(NutdbStmt) / 4), 0, acc);
  if (batch->n_tok) {
    k_hash<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)v.tok_type, batch->n_tok, 1, acc);
    k_hash<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)v.tok_start, batch->n_tok, 2, acc);
    k_hash<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)v.tok_end, batch->n_tok, 3, acc);
    k_hash<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)v.tok_kw, batch->n_tok, 4, acc);
  }
  if (batch->n_node) k_hash<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)v.node, batch->n_node, 5, acc);
  if (batch->n_err) k_hash<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)v.err, batch->n_err * (sizeof(NutdbError) / 4), 6, acc);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->hSmall.p, acc, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *out = *(const uint64_t*)ctx->hSmall.p;
  return NUTDB_OK;
}

int nutdb_gpu_batch_fetch_ext(NutdbBatch* batch) {
  if (!batch || !batch->impl) return NUTDB_E_ARG;
  NutdbCtx* ctx = (NutdbCtx*)batch->impl;
  if (!ctx->batch_live) return NUTDB_E_ARG;
  if (batch->ext || !batch->n_ext) return NUTDB_OK;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return NUTDB_E_CUDA;
  const int rc = fetch_ext(ctx);
  if (rc != NUTDB_OK) return rc;
  batch->ext = ctx->ext.data();
  return NUTDB_OK;
}

int nutdb_gpu_batch_device(const NutdbBatch* batch, NutdbBatchDevice* out) {
  if (!batch || !out || !batch->impl) return NUTDB_E_ARG;
  const NutdbCtx* ctx = (const NutdbCtx*)batch->impl;
  if (!ctx->batch_live) return NUTDB_E_ARG;
  *out = ctx->dev_view;
  return NUTDB_OK;
}

int nutdb_gpu_parse(NutdbCtx* ctx, const uint8_t* sql, uint64_t len, NutdbBatch* out) {
  const uint64_t off[2] = {0, len};
  static const uint8_t empty[1] = {0};
  return nutdb_gpu_parse_batch(ctx, sql ? sql : empty, off, 1, 0, out);
}

int nutdb_gpu_split_statements(NutdbCtx* ctx, const uint8_t* sql, uint64_t len, uint32_t flags, const uint64_t** stmt_off,
                               uint64_t* n_stmt) {
  if (!ctx) return NUTDB_E_CUDA;
  if (!stmt_off || !n_stmt || (len > 0 && !sql) || len >= 0x7FFFFFFFull) {
    ctx->err = "bad argument (null pointer, or 2^31 bytes or more)";
    return NUTDB_E_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ctx->launches = 0;
  ctx->recs.clear();
  ctx->ev_used = 0;
  const uint32_t n = (uint32_t)len;
  const uint32_t ntiles = (n + L2_TILE - 1) / L2_TILE;
  const size_t nchunks = (size_t)ntiles * (L2_TILE / 32);
  uint32_t* hS = (uint32_t*)ctx->hSmall.p;
  uint32_t* dS = (uint32_t*)ctx->small.p;
  uint64_t nsemi = 0;
  bool tail = false;
  if (n > 0) {
    const uint8_t* dText;
    if ((flags & NUTDB_F_DEVICE_INPUT) && (reinterpret_cast<uintptr_t>(sql) & 15u) == 0) {
      dText = sql;
    } else {
      ENSURE_DEV(text, (size_t)ntiles * L2_TILE + 16);
      CK(cudaMemcpyAsync(ctx->text.p, sql, n, (flags & NUTDB_F_DEVICE_INPUT) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
      dText = (const uint8_t*)ctx->text.p;
    }
    const size_t nwarps = (size_t)ntiles * L2_WARPS;
    ENSURE_DEV(bitmap, 4 * (nchunks + 1));
    ENSURE_DEV(localA, 4 * nwarps);
    ENSURE_DEV(tileA, 4 * (size_t)ntiles);
    ENSURE_DEV(entA, ntiles);
    ENSURE_DEV(winMasks, 4 * nchunks * L2_NMASK + 64);
    ENSURE_DEV(splitLocal, 8 * nwarps);
    ENSURE_DEV(splitTile, 8 * (size_t)ntiles);
    ENSURE_DEV(splitPref, 8 * (size_t)ntiles);
    // the whole buffer is ONE character stream: only byte 0 starts a "statement" for the context automaton
    CK(cudaMemsetAsync(ctx->bitmap.p, 0, 4 * (nchunks + 1), st));
    const uint32_t one = 1u;
    CK(cudaMemcpyAsync(ctx->bitmap.p, &one, 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(dS, 0, 64, st));
    const uint32_t* bm = (const uint32_t*)ctx->bitmap.p;
    LAUNCH("k_lex2_fn", k_lex2_fn<<<ntiles, L2_THREADS, 0, st>>>(dText, bm, n, ctx->dLex, ctx->dLex2, (uint32_t*)ctx->localA.p,
                                                                  (uint32_t*)ctx->tileA.p, (uint32_t*)ctx->winMasks.p, nchunks, nullptr));
    LAUNCH("k_scan_A", k_scan_vec8<<<1, SCAN_THREADS, 0, st>>>((const uint32_t*)ctx->tileA.p, (uint8_t*)ctx->entA.p, ntiles));
    LAUNCH("k_split_count", k_split<false><<<ntiles, L2_THREADS, 0, st>>>(
                                dText, bm, n, ctx->dLex, ctx->dLex2, (const uint32_t*)ctx->localA.p, (const uint8_t*)ctx->entA.p,
                                (uint32_t*)ctx->winMasks.p, nchunks, (uint2*)ctx->splitLocal.p, (uint2*)ctx->splitTile.p, nullptr,
                                nullptr));
    LAUNCH("k_scan_S", k_scan_tiles<U2AddOp><<<1, SCAN_THREADS, 0, st>>>((const uint2*)ctx->splitTile.p, (uint2*)ctx->splitPref.p,
                                                                         ntiles, (uint2*)(dS + 8)));
    CK(cudaMemcpyAsync(hS, dS, 64, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    nsemi = hS[8];
    ENSURE_DEV(splitOff, 8 * (nsemi + 2));
    ENSURE_HOST(hSplit, 8 * (nsemi + 3));
    LAUNCH("k_split_emit", k_split<true><<<ntiles, L2_THREADS, 0, st>>>(
                               dText, bm, n, ctx->dLex, ctx->dLex2, (const uint32_t*)ctx->localA.p, (const uint8_t*)ctx->entA.p,
                               (uint32_t*)ctx->winMasks.p, nchunks, (uint2*)ctx->splitLocal.p, nullptr,
                               (const uint2*)ctx->splitPref.p, (uint64_t*)ctx->splitOff.p));
    CK(cudaMemcpyAsync((uint64_t*)ctx->hSplit.p + 1, (uint64_t*)ctx->splitOff.p + 1, 8 * nsemi, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint64_t* off = (uint64_t*)ctx->hSplit.p;
    off[0] = 0;
    const uint64_t last = nsemi ? off[nsemi] : 0;
    if (last < n) {  // text after the last ';' is a statement of its own unless it is only whitespace
      k_tail_content<<<64, 256, 0, st>>>(dText, (uint32_t)last, n, dS + 12);
      ctx->launches++;
      CK(cudaMemcpyAsync(hS, dS, 64, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      tail = hS[12] != 0;
    }
    if (tail) off[nsemi + 1] = n;
  } else {
    ENSURE_HOST(hSplit, 64);
    ((uint64_t*)ctx->hSplit.p)[0] = 0;
  }
  CK(cudaGetLastError());
  *stmt_off = (const uint64_t*)ctx->hSplit.p;
  *n_stmt = nsemi + (tail ? 1 : 0);
  return NUTDB_OK;
}

int nutdb_gpu_last_timing(const NutdbCtx* ctx, float ms[5]) {
  if (!ctx || !ms) return NUTDB_E_ARG;
  for (int i = 0; i < 5; i++) ms[i] = ctx->ms[i];
  return NUTDB_OK;
}

int nutdb_gpu_last_launches(const NutdbCtx* ctx) { return ctx ? ctx->launches : 0; }

// test hook (not in the public header): the first statements of the last batch that the table-driven parser declined
int nutdb_gpu_debug_slow_list(NutdbCtx* ctx, uint32_t* out, uint32_t cap) {
  if (!ctx || !ctx->slowList2.p) return 0;
  const uint32_t k = std::min<uint32_t>(cap, ctx->n_slow);
  cudaSetDevice(ctx->device);
  if (k && cudaMemcpy(out, ctx->slowList2.p, 4 * (size_t)k, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)k;
}

// test hook (not in the public header): lex every batch with k_lex3 (look-back scans) instead of k_lex4 (ranges)
void nutdb_gpu_debug_force_lookback(NutdbCtx* ctx, int on) {
  if (ctx) ctx->force_lookback = on != 0;
}
int nutdb_gpu_debug_last_lookback(const NutdbCtx* ctx) { return ctx && ctx->last_lookback ? 1 : 0; }

// profiling hook (not in the public header): 24 clock64 stamps per tile of the last k_lex3 launch
int nutdb_gpu_debug_timing(NutdbCtx* ctx, int enable, unsigned long long* out, uint32_t cap_tiles) {
  if (!ctx) return -1;
  ctx->debug_timing = enable != 0;
  if (!out || !ctx->dbgTim.p) return 0;
  const uint32_t nt = std::min(cap_tiles, ctx->dbg_ntiles);
  cudaSetDevice(ctx->device);
  if (cudaMemcpy(out, ctx->dbgTim.p, 24 * 8 * (size_t)nt, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)nt;
}

// test hook (not in the public header): per-tile carries of the single-pass lexer's two look-back scans
int nutdb_gpu_debug_tiles(NutdbCtx* ctx, int enable, uint32_t* out, uint32_t cap_tiles) {
  if (!ctx) return -1;
  ctx->debug_tiles = enable != 0;
  if (!out || !ctx->dbgTiles.p) return 0;
  const uint32_t nt = std::min(cap_tiles, ctx->dbg_ntiles);
  cudaSetDevice(ctx->device);
  if (cudaMemcpy(out, ctx->dbgTiles.p, 16 * (size_t)nt, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)nt;
}

void nutdb_gpu_set_profiling(NutdbCtx* ctx, int on) {
  if (ctx) ctx->profiling = on != 0;
}
int nutdb_gpu_kernel_timing(const NutdbCtx* ctx, int i, const char** name, float* ms) {
  if (!ctx) return 0;
  if (i >= 0 && (size_t)i < ctx->kernel_ms.size()) {
    if (name) *name = ctx->kernel_ms[i].first;
    if (ms) *ms = ctx->kernel_ms[i].second;
  }
  return (int)ctx->kernel_ms.size();
}
uint64_t nutdb_gpu_last_slow_statements(const NutdbCtx* ctx) { return ctx ? ctx->n_slow : 0; }
uint64_t nutdb_gpu_last_wide_statements(const NutdbCtx* ctx) { return ctx ? ctx->n_wide : 0; }
uint64_t nutdb_gpu_last_exact_lexed_statements(const NutdbCtx* ctx) { return ctx ? ctx->n_punt : 0; }
void* nutdb_gpu_ctx_stream(const NutdbCtx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"
