// lex2_kernels.cuh -- kernels of the warp-cooperative lexer (logic in lex2_core.cuh).
// Included by nutdb_gpu.cu after the scan helpers and DevSink.
//
//   k_lex2_fn      per warp (1 KB): transition function of the context automaton; per-tile aggregate
//   k_lex2_walk<0> per warp: number of tokens + string / statement carries; flags statements for the exact path
//   k_lex2_walk<1> per warp: emits tokens at their final index
//   k_lex_exact<0/1> one thread per flagged statement: the exact walker of lex_core.cuh (count / emit)
#pragma once

#define L2_THREADS 256
// resident CTAs per SM the walk kernels are compiled for (register cap): measured best, counting pass 5 (48
// registers), emitting pass 4 (64)
#ifndef L2_COUNT_MINBLOCKS
#define L2_COUNT_MINBLOCKS 5
#endif
#ifndef L2_EMIT_MINBLOCKS
#define L2_EMIT_MINBLOCKS 4
#endif
#ifndef L2_FN_MINBLOCKS
#define L2_FN_MINBLOCKS 8
#endif
#define L2_WARPS (L2_THREADS / 32)
#define L2_SEG 1024                     // bytes per warp
#define L2_TILE (L2_WARPS * L2_SEG)     // bytes per block

#define L2_HALO 32  // bytes in front of and behind the tile that are staged too (look-back / look-ahead of edge windows)
struct alignas(16) Lex2Shared {
  alignas(16) uint32_t text[(L2_TILE + 2 * L2_HALO) / 4];  // [tile_begin - 32, tile_end + 32), filled with 16-byte stores
  uint32_t bm[L2_TILE / 32];
  LexTables T;
  nlex2::Lex2Tables K;
};

// bytes outside the CTA's tile (look-back / look-ahead across a tile edge): rare, kept out of line
__device__ __noinline__ uint8_t far_byte(const uint8_t* text, uint32_t p, uint32_t n) { return p < n ? text[p] : (uint8_t)0; }

struct Tile2Src {
  const uint8_t* text;
  const uint8_t* sm;
  uint32_t tile_begin, n;
  // sm points at the tile's first byte; the halo lies at sm[-32..-1] and sm[L2_TILE..L2_TILE+31]
  __device__ __forceinline__ uint8_t byte(uint32_t p) const {
    const uint32_t r = p - tile_begin + L2_HALO;
    if (r < L2_TILE + 2 * L2_HALO) return sm[(int)r - L2_HALO];
    return far_byte(text, p, n);
  }
  // [p, p + len) as one contiguous run of shared memory, or nullptr if it leaves the staged range
  __device__ __forceinline__ const uint8_t* span(uint32_t p, uint32_t len) const {
    const uint32_t r = p - tile_begin + L2_HALO;
    return (r < L2_TILE + 2 * L2_HALO && r + len <= L2_TILE + 2 * L2_HALO) ? sm + ((int)r - L2_HALO) : nullptr;
  }
};

__device__ __forceinline__ void stage_tile2(const uint8_t* text, const uint32_t* bitmap, uint32_t tile_begin, uint32_t n,
                                            Lex2Shared& S, const LexTables* gT, const nlex2::Lex2Tables* gK) {
  stage_tables(gT, &S.T);
  {
    const uint32_t* a = reinterpret_cast<const uint32_t*>(gK);
    uint32_t* b = reinterpret_cast<uint32_t*>(&S.K);
    for (uint32_t i = threadIdx.x; i < sizeof(nlex2::Lex2Tables) / 4; i += L2_THREADS) b[i] = a[i];
  }
  const uint4* src = reinterpret_cast<const uint4*>(text + tile_begin);
  uint4* dst = reinterpret_cast<uint4*>(S.text) + L2_HALO / 16;
#pragma unroll
  for (uint32_t k = threadIdx.x; k < L2_TILE / 16; k += L2_THREADS) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (tile_begin + 16u * k < n) v = __ldg(src + k);
    dst[k] = v;
  }
  if (threadIdx.x < 2 * L2_HALO / 16) {  // the halo: two 16-byte pieces on either side (zero outside the batch)
    const int k = threadIdx.x < L2_HALO / 16 ? (int)threadIdx.x - L2_HALO / 16 : (int)(L2_TILE / 16) + (int)threadIdx.x - L2_HALO / 16;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const long long p = (long long)tile_begin + 16ll * k;
    if (p >= 0 && p < (long long)n) v = __ldg(src + k);
    dst[k] = v;
  }
  S.bm[threadIdx.x] = bitmap[(tile_begin >> 5) + threadIdx.x];
}

// boundary word of the window at `base` (multiple of 32): tile-local from shared memory, else global
__device__ __forceinline__ uint32_t bnd_word(const Lex2Shared& S, const uint32_t* bitmap, uint32_t tile_begin,
                                             uint32_t base) {
  const uint32_t r = base - tile_begin;
  return r < L2_TILE ? S.bm[r >> 5] : bitmap[base >> 5];
}

// ---- per-thread window set-up shared by the three kernels -------------------------------------------------
// Lane j of a warp owns window j of the warp's 1 KB block.  stage 1: the warp classifies window jj (one byte
// per lane), ballots give its class masks, lane jj keeps them.
template <bool All>
__device__ __forceinline__ void build_masks(const Lex2Shared& S, const Tile2Src& src, uint32_t blk, uint32_t n, int lane,
                                            nlex2::Win& w) {
  const uint32_t full = 0xFFFFFFFFu;
  w.sq = w.dq = w.bt = w.nl = w.bs = w.dash = w.slash = w.star = w.L = w.D = w.DOT = w.OP = w.P = w.WS = w.IE = w.NE = w.u = 0u;
#pragma unroll 4
  for (int jj = 0; jj < 32; jj++) {
    const uint32_t pos = blk + 32u * (uint32_t)jj + (uint32_t)lane;
    const uint16_t k = pos < n ? S.K.cls[src.sm[pos - src.tile_begin]] : (uint16_t)0;  // always inside the tile
    const bool mine = lane == jj;
    uint32_t m;
    m = __ballot_sync(full, k & nlex2::K_SQ); if (mine) w.sq = m;
    m = __ballot_sync(full, k & nlex2::K_DQ); if (mine) w.dq = m;
    m = __ballot_sync(full, k & nlex2::K_BT); if (mine) w.bt = m;
    m = __ballot_sync(full, k & nlex2::K_NL); if (mine) w.nl = m;
    m = __ballot_sync(full, k & nlex2::K_BS); if (mine) w.bs = m;
    m = __ballot_sync(full, k & nlex2::K_DASH); if (mine) w.dash = m;
    m = __ballot_sync(full, k & nlex2::K_SLASH); if (mine) w.slash = m;
    m = __ballot_sync(full, k & nlex2::K_STAR); if (mine) w.star = m;
    if (All) {
      m = __ballot_sync(full, k & nlex2::K_L); if (mine) w.L = m;
      m = __ballot_sync(full, k & nlex2::K_D); if (mine) w.D = m;
      m = __ballot_sync(full, k & nlex2::K_DOT); if (mine) w.DOT = m;
      m = __ballot_sync(full, k & nlex2::K_OP); if (mine) w.OP = m;
      m = __ballot_sync(full, k & nlex2::K_P); if (mine) w.P = m;
      m = __ballot_sync(full, k & nlex2::K_WS); if (mine) w.WS = m;
    }
  }
}

// backslash parity and previous byte in front of position pos (walks back over the run; never across a statement start)
__device__ __forceinline__ void entry_esc(const Tile2Src& src, const uint32_t* bitmap, uint32_t pos, uint8_t& prev,
                                          uint8_t& esc) {
  prev = 0;
  esc = 0;
  if (pos == 0) return;
  prev = src.byte(pos - 1);
  if ((bitmap[pos >> 5] >> (pos & 31u)) & 1u) return;  // a statement starts here
  uint32_t nrun = 0, p = pos;
  while (p > 0 && src.byte(p - 1) == '\\') {
    nrun++;
    p--;
    if ((bitmap[p >> 5] >> (p & 31u)) & 1u) break;
  }
  esc = (uint8_t)(nrun & 1u);
}

// stage 1, per thread: the 32 bytes of the lane's own window -> 32 class words (16 bits each) through the
// shared-memory class table, then a SWAR bit-matrix transposition (two 16x16 blocks side by side in 16
// registers, 4 butterfly stages) turns "16 class bits per byte" into "32 byte bits per class".
__device__ __forceinline__ void build_masks_transpose(const Lex2Shared& S, uint32_t tile_begin, uint32_t base, uint32_t valid,
                                                      nlex2::Win& w) {
  const uint4* wp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(S.text) + L2_HALO + (base - tile_begin));
  const uint4 q0 = wp[0], q1 = wp[1];
  const uint32_t lo[4] = {q0.x, q0.y, q0.z, q0.w}, hi[4] = {q1.x, q1.y, q1.z, q1.w};
  uint32_t A[16];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int b = 0; b < 4; b++)
      A[4 * i + b] = (uint32_t)S.K.cls[(lo[i] >> (8 * b)) & 0xFFu] | ((uint32_t)S.K.cls[(hi[i] >> (8 * b)) & 0xFFu] << 16);
  uint32_t m = 0x00FF00FFu;
#pragma unroll
  for (int j = 8; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 16; k = (k + j + 1) & ~j) {
      const uint32_t t = ((A[k] >> j) ^ A[k + j]) & m;
      A[k + j] ^= t;
      A[k] ^= t << j;
    }
  }
  w.sq = A[0] & valid;
  w.dq = A[1] & valid;
  w.bt = A[2] & valid;
  w.nl = A[3] & valid;
  w.bs = A[4] & valid;
  w.dash = A[5] & valid;
  w.slash = A[6] & valid;
  w.star = A[7] & valid;
  w.L = A[8] & valid;
  w.D = A[9] & valid;
  w.DOT = A[10] & valid;
  w.OP = A[11] & valid;
  w.P = A[12] & valid;
  w.WS = A[13] & valid;
  w.IE = A[14] & valid;
  w.NE = A[15] & valid;
  w.u = 0u;
}

// class masks of every window, computed once by k_lex2_fn: 14 arrays of nwin words (structure of arrays)
#define L2_NMASK 16
__device__ __forceinline__ void store_masks(uint32_t* __restrict__ g, size_t stride, uint32_t win, const nlex2::Win& w) {
  g[0 * stride + win] = w.sq;
  g[1 * stride + win] = w.dq;
  g[2 * stride + win] = w.bt;
  g[3 * stride + win] = w.nl;
  g[4 * stride + win] = w.bs;
  g[5 * stride + win] = w.dash;
  g[6 * stride + win] = w.slash;
  g[7 * stride + win] = w.star;
  g[8 * stride + win] = w.L;
  g[9 * stride + win] = w.D;
  g[10 * stride + win] = w.DOT;
  g[11 * stride + win] = w.OP;
  g[12 * stride + win] = w.P;
  g[13 * stride + win] = w.WS;
  g[14 * stride + win] = w.IE;
  g[15 * stride + win] = w.NE;
}
__device__ __forceinline__ void load_masks(const uint32_t* __restrict__ g, size_t stride, uint32_t win, nlex2::Win& w) {
  w.sq = g[0 * stride + win];
  w.dq = g[1 * stride + win];
  w.bt = g[2 * stride + win];
  w.nl = g[3 * stride + win];
  w.bs = g[4 * stride + win];
  w.dash = g[5 * stride + win];
  w.slash = g[6 * stride + win];
  w.star = g[7 * stride + win];
  w.L = g[8 * stride + win];
  w.D = g[9 * stride + win];
  w.DOT = g[10 * stride + win];
  w.OP = g[11 * stride + win];
  w.P = g[12 * stride + win];
  w.WS = g[13 * stride + win];
  w.IE = g[14 * stride + win];
  w.NE = g[15 * stride + win];
  w.u = 0u;
}

struct WinSetup {
  nlex2::Win w;
  nlex2::Next nx;
  nlex2::Events ev;
  uint32_t base, escm;
  uint8_t prev_byte;
};

// everything a thread knows about its window before any carry: masks, neighbours, escapes, events
template <bool Build, bool Virt>
__device__ __forceinline__ void setup_window(const Lex2Shared& S, const Tile2Src& src, const uint32_t* bitmap,
                                             uint32_t tile_begin, uint32_t blk, uint32_t n, int lane, uint32_t* gmask,
                                             size_t mstride, WinSetup& u) {
  const uint32_t full = 0xFFFFFFFFu;
  u.base = blk + 32u * (uint32_t)lane;
  nlex2::Win& w = u.w;
  w.valid = u.base + 32u <= n ? full : (n > u.base ? ((1u << (n - u.base)) - 1u) : 0u);
  if (Build) {
    build_masks_transpose(S, tile_begin, u.base, w.valid, w);
    store_masks(gmask, mstride, u.base >> 5, w);
  } else {
    load_masks(gmask, mstride, u.base >> 5, w);
  }
  uint32_t bnd = u.base < n ? (bnd_word(S, bitmap, tile_begin, u.base) & w.valid) : 0u;
  if (Virt && n >= u.base && n - u.base < 32u) bnd |= 1u << (n - u.base);  // the batch end terminates the last statement
  w.bnd = bnd;
  // next window's first byte
  const uint32_t p = u.base + 32u;
  if (p >= n) {
    u.nx.byte = 0;
    u.nx.bnd = 1;
    u.nx.cls = 0;
  } else {
    u.nx.byte = src.byte(p);
    u.nx.bnd = (uint8_t)(bnd_word(S, bitmap, tile_begin, p) & 1u);
    u.nx.cls = S.K.cls[u.nx.byte];
  }
  // previous byte and backslash parity in front of the window
  const uint32_t bs_prev = __shfl_up_sync(full, w.bs, 1);
  uint8_t esc_in;
  if (lane == 0) {
    entry_esc(src, bitmap, u.base < n ? u.base : 0u, u.prev_byte, esc_in);
  } else {
    u.prev_byte = u.base <= n && u.base > 0 ? src.byte(u.base - 1u) : (uint8_t)0;
    const int run = nlex2::clz32(~bs_prev);
    esc_in = (uint8_t)(run >= 32 ? 0 : (run & 1));  // (a window of 32 backslashes: its statement is flagged, see below)
  }
  u.escm = nlex2::esc_mask32(w.bs, esc_in) & ~w.bnd;
  u.ev = nlex2::make_events(w, u.escm, u.prev_byte);
}

__device__ __forceinline__ uint32_t warp_scan_vec8(uint32_t f, int lane, uint32_t& excl) {
  const uint32_t full = 0xFFFFFFFFu;
  uint32_t v = f;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(full, v, d);
    if (lane >= d) v = vec8_then(o, v);
  }
  excl = __shfl_up_sync(full, v, 1);
  if (lane == 0) excl = NUTDB_VEC8_ID;
  return v;
}

__device__ __forceinline__ uint32_t window_fn(const LexTables& T, const WinSetup& u) {
  if (u.ev.all) return nlex2::ctx_window_fn(T, u.w, u.ev, NUTDB_VEC8_ID);
  return vec8_then_row(NUTDB_VEC8_ID, T.a_row[EV_OTHER][0], T.a_row[EV_OTHER][1]);
}

__global__ void __launch_bounds__(L2_THREADS, L2_FN_MINBLOCKS) k_lex2_fn(const uint8_t* __restrict__ text, const uint32_t* __restrict__ bitmap,
                                                        uint32_t n, const LexTables* __restrict__ gT,
                                                        const nlex2::Lex2Tables* __restrict__ gK,
                                                        uint32_t* __restrict__ localA, uint32_t* __restrict__ tileA,
                                                        uint32_t* __restrict__ gmask, size_t mstride,
                                                        uint32_t* __restrict__ winfn /* may be null */) {
  __shared__ Lex2Shared S;
  __shared__ uint32_t wfn[L2_WARPS];
  const uint32_t tile_begin = blockIdx.x * L2_TILE;
  stage_tile2(text, bitmap, tile_begin, n, S, gT, gK);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Tile2Src src{text, reinterpret_cast<const uint8_t*>(S.text) + L2_HALO, tile_begin, n};
  const uint32_t blk = tile_begin + (uint32_t)warp * L2_SEG;
  uint32_t run = NUTDB_VEC8_ID;
  if (blk < n) {
    WinSetup u;
    setup_window<true, false>(S, src, bitmap, tile_begin, blk, n, lane, gmask, mstride, u);
    const uint32_t f = u.base < n ? window_fn(S.T, u) : NUTDB_VEC8_ID;
    uint32_t excl;
    run = __shfl_sync(0xFFFFFFFFu, warp_scan_vec8(f, lane, excl), 31);
    // the composed function of the windows before this one in the warp: the counting pass applies it to the warp's
    // entry state instead of walking the events a second time
    if (winfn && u.base < n) winfn[u.base >> 5] = excl;
  }
  if (lane == 0) wfn[warp] = run;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = NUTDB_VEC8_ID;
    for (int i = 0; i < L2_WARPS; i++) {
      localA[blockIdx.x * L2_WARPS + i] = acc;
      acc = vec8_then(acc, wfn[i]);
    }
    tileA[blockIdx.x] = acc;
  }
}

struct Lex2Out {
  uint8_t* type;
  uint32_t* start;
  uint32_t* end;
  uint8_t* kw;
  uint32_t cap;
  uint32_t* stmt_tok_begin;
  uint32_t* stmt_tok_end;
  const uint32_t* off32;
  uint32_t nstmt;
  uint32_t* punt_flag;   // per statement
  uint32_t* punt_count;  // [0] flagged statements, [1] bound on their tokens
  const uint32_t* first_stmt;  // per 32-byte window: smallest non-empty statement starting in it
  uint32_t nbytes;
  __device__ uint32_t find_stmt(uint32_t pos) const {  // the statement containing byte `pos`
    uint32_t lo = 0, hi = nstmt;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (off32[mid] > pos) hi = mid;
      else lo = mid + 1;
    }
    return lo - 1;
  }
  __device__ void punt(uint32_t pos) const {
    const uint32_t s = find_stmt(pos);
    if (atomicExch(&punt_flag[s], 1u) == 0u) {
      atomicAdd(punt_count, 1u);
      // a bound on the tokens the exact lexer can produce for it (a token per byte + EOF): sizes the extra region
      atomicAdd(punt_count + 1, off32[s + 1] - off32[s] + 1u);
    }
  }
  // sink interface of nlex2::win_emit
  __device__ __forceinline__ void token(uint32_t i, uint8_t t, uint32_t s, uint32_t e, uint8_t k) const {
    if (i < cap) {
      type[i] = t;
      start[i] = s;
      end[i] = e;
      kw[i] = k;
    }
  }
  // A statement starts at byte `pos` and its first token has index `first`.  first_stmt[pos / 32] is the
  // smallest non-empty statement starting in that window (k_prep), so the statement is found by stepping over
  // the few statements of the window (and over empty ones, which share their offset with their successor).
  // The same lane closes the token range of the previous non-empty statement: its EOF token is first - 1.
  __device__ __forceinline__ void stmt_begin(uint32_t pos, uint32_t first) const {
    uint32_t c = first_stmt[pos >> 5];
    while (off32[c] != pos || off32[c + 1] == pos) c++;
    // (a statement flagged for the exact lexer gets its token range from k_lex_exact, which runs beside this pass)
    if (punt_flag[c] == 0u) stmt_tok_begin[c] = first;
    uint32_t p = c;
    while (p > 0) {
      p--;
      if (off32[p + 1] != off32[p]) {
        if (punt_flag[p] == 0u) stmt_tok_end[p] = first;
        break;
      }
    }
  }
  // only the last statement of the batch has no successor to close its range
  __device__ __forceinline__ void stmt_end(uint32_t pos, uint32_t endi) const {
    if (pos + 1u != nbytes) return;
    uint32_t p = nstmt;
    while (p > 0) {
      p--;
      if (off32[p + 1] != off32[p]) {
        if (punt_flag[p] == 0u) stmt_tok_end[p] = endi;
        break;
      }
    }
  }
};

// Emit = false: token-end mask per window (whas) + per-warp carries + flags; Emit = true: tokens.
template <bool Emit>
__global__ void __launch_bounds__(L2_THREADS, Emit ? L2_EMIT_MINBLOCKS : L2_COUNT_MINBLOCKS) k_lex2_walk(const uint8_t* __restrict__ text,
                                                          const uint32_t* __restrict__ bitmap, uint32_t n,
                                                          const LexTables* __restrict__ gT,
                                                          const nlex2::Lex2Tables* __restrict__ gK,
                                                          const uint32_t* __restrict__ localA,
                                                          const uint8_t* __restrict__ tileEntA, uint4* __restrict__ localC,
                                                          uint4* __restrict__ tileC, const uint4* __restrict__ tilePrefC,
                                                          uint32_t* __restrict__ whas, uint32_t* __restrict__ wctx,
                                                          uint32_t* __restrict__ gmask, size_t mstride, Lex2Out out,
                                                          const uint32_t* __restrict__ winfn) {
  __shared__ Lex2Shared S;
  __shared__ uint4 wsum[L2_WARPS];
  const uint32_t tile_begin = blockIdx.x * L2_TILE;
  stage_tile2(text, bitmap, tile_begin, n, S, gT, gK);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t full = 0xFFFFFFFFu;
  Tile2Src src{text, reinterpret_cast<const uint8_t*>(S.text) + L2_HALO, tile_begin, n};
  const uint32_t blk = tile_begin + (uint32_t)warp * L2_SEG;
  const uint32_t widx = blockIdx.x * L2_WARPS + warp;
  CSum wtotal = csum_identity();
  if (blk < n) {
    WinSetup u;
    setup_window<false, true>(S, src, bitmap, tile_begin, blk, n, lane, gmask, mstride, u);
    const bool live = u.base < n;
    // entry state of every window: scan of the window functions on top of the warp's entry state
    // (the counting pass stores it, one byte per window, for the emitting pass)
    const uint8_t s_warp = (uint8_t)vec8_apply(localA[widx], tileEntA[blockIdx.x]);
    uint8_t s_in;
    if (!Emit) {
      const uint32_t excl = live ? winfn[u.base >> 5] : NUTDB_VEC8_ID;  // from k_lex2_fn
      s_in = (uint8_t)vec8_apply(excl, s_warp);
    } else {
      s_in = (uint8_t)A_C;
    }
    // the concrete walk of the window's events is done once, by the counting pass; the emitting pass reads what it
    // needs of the result (5 words per window)
    nlex2::WinCtx o;
    if (live) {
      if (!Emit) {
        nlex2::ctx_window(u.w, u.ev, u.base, u.nx, s_in, u.prev_byte, o);
        const nlex2::WinCtxPacked pk = nlex2::ctx_pack(o, u.base);
#pragma unroll
        for (int q = 0; q < 5; q++) wctx[(size_t)q * mstride + (u.base >> 5)] = pk.w[q];
      } else {
        nlex2::WinCtxPacked pk;
#pragma unroll
        for (int q = 0; q < 5; q++) pk.w[q] = wctx[(size_t)q * mstride + (u.base >> 5)];
        nlex2::ctx_unpack(pk, u.base, o);
      }
    }
    o.escm = u.escm;
    // history: code-token class masks of the previous window (lane 0: the 32 bytes in front of the block, whose
    // raw classes are exact when the block is entered in code)
    nlex2::Hist h;
    {
      uint32_t rL = 0, rD = 0, rDOT = 0, rOP = 0;
      if (blk >= 32u) {
        const uint16_t k = S.K.cls[src.byte(blk - 32u + (uint32_t)lane)];
        rL = __ballot_sync(full, k & nlex2::K_L);
        rD = __ballot_sync(full, k & nlex2::K_D);
        rDOT = __ballot_sync(full, k & nlex2::K_DOT);
        rOP = __ballot_sync(full, k & nlex2::K_OP);
      }
      const bool code_entry = s_warp <= A_CX;
      h.L = __shfl_up_sync(full, u.w.L & o.ct, 1);
      h.D = __shfl_up_sync(full, u.w.D & o.ct, 1);
      h.DOT = __shfl_up_sync(full, u.w.DOT & o.ct, 1);
      h.OP = __shfl_up_sync(full, u.w.OP & o.ct, 1);
      h.bnd = __shfl_up_sync(full, u.w.bnd, 1);
      if (lane == 0) {
        h.L = code_entry ? rL : 0u;
        h.D = code_entry ? rD : 0u;
        h.DOT = code_entry ? rDOT : 0u;
        h.OP = code_entry ? rOP : 0u;
        h.bnd = blk >= 32u ? bitmap[(blk - 32u) >> 5] : 0u;
      }
    }
    nlex2::StrCarry sc_in;
    uint32_t stmt_in = 0, index = 0, ntok = 0;
    if (!Emit) {
      uint32_t bad = 0;
      if (live) {
        const uint32_t has = nlex2::win_has_mask(S.T, src, u.w, o, h, u.nx, u.base, u.prev_byte, bad);
        ntok = (uint32_t)__popc(has) + (uint32_t)__popc(nlex2::win_eof_mask(u.w, u.nx));
        if (u.w.bs == 0xFFFFFFFFu) bad |= 1u;  // backslash run longer than a window: parity not tracked
        whas[u.base >> 5] = has;
        uint32_t bb = bad;
        while (bb) {
          const int i = __ffs((int)bb) - 1;
          bb &= bb - 1;
          out.punt(u.base + (uint32_t)i);
        }
        bb = o.bad_prev;
        while (bb) {
          const int i = __ffs((int)bb) - 1;
          bb &= bb - 1;
          out.punt(u.base + (uint32_t)i - 1u);
        }
        // the batch ends exactly on a window boundary: no window carries the virtual end-of-batch statement start
        if (u.base + 32u == n && (o.s_out == A_SQ || o.s_out == A_DQ || o.s_out == A_BT || o.s_out == A_BC0 || o.s_out == A_BC))
          out.punt(n - 1u);
      }
      // warp totals: count, open string, last statement start (ordered reductions over the lanes)
      uint32_t cnt = ntok, lb = o.last_bnd1;
      uint32_t sc_open = o.sc.has_open, sc_esc = o.sc.esc, sc_pos = o.sc.open_pos;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t c2 = __shfl_up_sync(full, cnt, d), l2 = __shfl_up_sync(full, lb, d);
        const uint32_t o2 = __shfl_up_sync(full, sc_open, d), e2 = __shfl_up_sync(full, sc_esc, d),
                       p2 = __shfl_up_sync(full, sc_pos, d);
        if (lane >= d) {
          cnt += c2;
          lb = max(lb, l2);
          if (!sc_open) {  // str_then(earlier, mine)
            sc_open = o2;
            sc_esc |= e2;
            sc_pos = p2;
          }
        }
      }
      if (lane == 31) {
        wtotal.count = cnt;
        wtotal.nseg = lb ? 1u : 0u;
        wtotal.stmt_start = lb ? lb - 1u : 0u;
        wtotal.has_tok = (uint8_t)sc_open;
        wtotal.tok_start = sc_pos;
        wtotal.escaped = (uint8_t)sc_esc;
        wsum[warp] = csum_pack(wtotal);
      }
    } else {
      const CSum pre = csum_unpack(CSumOp::then(tilePrefC[blockIdx.x], localC[widx]));
      // exclusive scans over the lanes, seeded with the warp's carry-in
      const uint32_t has = live ? whas[u.base >> 5] : 0u;
      uint32_t cnt = live ? (uint32_t)__popc(has) + (uint32_t)__popc(nlex2::win_eof_mask(u.w, u.nx)) : 0u, lb = o.last_bnd1;
      uint32_t sc_open = o.sc.has_open, sc_esc = o.sc.esc, sc_pos = o.sc.open_pos;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t c2 = __shfl_up_sync(full, cnt, d), l2 = __shfl_up_sync(full, lb, d);
        const uint32_t o2 = __shfl_up_sync(full, sc_open, d), e2 = __shfl_up_sync(full, sc_esc, d),
                       p2 = __shfl_up_sync(full, sc_pos, d);
        if (lane >= d) {
          cnt += c2;
          lb = max(lb, l2);
          if (!sc_open) {
            sc_open = o2;
            sc_esc |= e2;
            sc_pos = p2;
          }
        }
      }
      uint32_t xcnt = __shfl_up_sync(full, cnt, 1), xlb = __shfl_up_sync(full, lb, 1);
      uint32_t xo = __shfl_up_sync(full, sc_open, 1), xe = __shfl_up_sync(full, sc_esc, 1), xp = __shfl_up_sync(full, sc_pos, 1);
      if (lane == 0) {
        xcnt = 0;
        xlb = 0;
        xo = 0;
        xe = 0;
        xp = 0;
      }
      index = pre.count + xcnt;
      stmt_in = xlb ? xlb - 1u : pre.stmt_start;
      if (xo) {
        sc_in.has_open = 1;
        sc_in.esc = (uint8_t)xe;
        sc_in.open_pos = xp;
      } else {
        sc_in.has_open = 0;
        sc_in.esc = (uint8_t)(pre.escaped | xe);
        sc_in.open_pos = pre.tok_start;
      }
      if (live) nlex2::win_emit(S.T, src, out, u.w, o, h, u.nx, u.base, u.prev_byte, sc_in, stmt_in, index, has);
    }
  } else if (!Emit) {
    if (lane == 31) wsum[warp] = csum_pack(wtotal);
  }
  if (!Emit) {
    __syncthreads();
    if (threadIdx.x == 0) {
      uint4 acc = CSumOp::identity();
      for (int i = 0; i < L2_WARPS; i++) {
        localC[blockIdx.x * L2_WARPS + i] = acc;
        acc = CSumOp::then(acc, wsum[i]);
      }
      tileC[blockIdx.x] = acc;
    }
  }
}

// ---- statement splitter: every ';' in code context ends a statement -----------------
// bit i: byte i of the thread's window is ';'
__device__ __forceinline__ uint32_t semicolon_mask(const Lex2Shared& S, uint32_t tile_begin, uint32_t base, uint32_t valid) {
  const uint4* wp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(S.text) + L2_HALO + (base - tile_begin));
  const uint4 q0 = wp[0], q1 = wp[1];
  const uint32_t v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t x = v[i] ^ 0x3B3B3B3Bu;                                   // zero byte <=> ';'
    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);  // 0x80 in every zero byte
    m |= (((z >> 7) * 0x00204081u) >> 21 & 0xFu) << (4 * i);                 // gather the four flags
  }
  return m & valid;
}

// Emit = false: number of code-context ';' per warp block / tile; Emit = true: offsets[1 + k] = position after the k-th
template <bool Emit>
__global__ void __launch_bounds__(L2_THREADS) k_split(const uint8_t* __restrict__ text, const uint32_t* __restrict__ bitmap,
                                                      uint32_t n, const LexTables* __restrict__ gT,
                                                      const nlex2::Lex2Tables* __restrict__ gK,
                                                      const uint32_t* __restrict__ localA, const uint8_t* __restrict__ tileEntA,
                                                      uint32_t* __restrict__ gmask, size_t mstride, uint2* __restrict__ localS,
                                                      uint2* __restrict__ tileS, const uint2* __restrict__ tilePrefS,
                                                      uint64_t* __restrict__ offsets) {
  __shared__ Lex2Shared S;
  __shared__ uint32_t wsum[L2_WARPS];
  const uint32_t tile_begin = blockIdx.x * L2_TILE;
  stage_tile2(text, bitmap, tile_begin, n, S, gT, gK);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t full = 0xFFFFFFFFu;
  Tile2Src src{text, reinterpret_cast<const uint8_t*>(S.text) + L2_HALO, tile_begin, n};
  const uint32_t blk = tile_begin + (uint32_t)warp * L2_SEG;
  const uint32_t widx = blockIdx.x * L2_WARPS + warp;
  uint32_t total = 0;
  if (blk < n) {
    WinSetup u;
    setup_window<false, false>(S, src, bitmap, tile_begin, blk, n, lane, gmask, mstride, u);
    const bool live = u.base < n;
    const uint32_t f = live ? window_fn(S.T, u) : NUTDB_VEC8_ID;
    uint32_t excl;
    warp_scan_vec8(f, lane, excl);
    const uint8_t s_warp = (uint8_t)vec8_apply(localA[widx], tileEntA[blockIdx.x]);
    const uint8_t s_in = (uint8_t)vec8_apply(excl, s_warp);
    nlex2::WinCtx o;
    if (live) nlex2::ctx_window(u.w, u.ev, u.base, u.nx, s_in, u.prev_byte, o);
    const uint32_t semis = live ? (o.ct & semicolon_mask(S, tile_begin, u.base, u.w.valid)) : 0u;
    uint32_t cnt = (uint32_t)__popc(semis);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t c2 = __shfl_up_sync(full, cnt, d);
      if (lane >= d) cnt += c2;
    }
    total = __shfl_sync(full, cnt, 31);
    if (Emit) {
      uint32_t idx = tilePrefS[blockIdx.x].x + localS[widx].x + cnt - (uint32_t)__popc(semis);
      uint32_t todo = semis;
      while (todo) {
        const int i = __ffs((int)todo) - 1;
        todo &= todo - 1;
        offsets[1u + idx++] = (uint64_t)u.base + (uint64_t)i + 1ull;
      }
    }
  }
  if (!Emit) {
    if (lane == 0) wsum[warp] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t acc = 0;
      for (int i = 0; i < L2_WARPS; i++) {
        localS[blockIdx.x * L2_WARPS + i] = make_uint2(acc, 0u);
        acc += wsum[i];
      }
      tileS[blockIdx.x] = make_uint2(acc, 0u);
    }
  }
}

// is there anything but whitespace in [from, n)?
__global__ void k_tail_content(const uint8_t* __restrict__ text, uint32_t from, uint32_t n, uint32_t* __restrict__ flag) {
  for (uint32_t p = from + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const uint8_t b = text[p];
    if (!(b == ' ' || b == '\t' || b == '\n' || b == '\r')) {
      atomicOr(flag, 1u);
      return;
    }
  }
}

// ---- exact path for flagged statements: one thread per statement runs the walker of lex_core.cuh ----
struct StmtSrc {
  const uint8_t* text;
  uint32_t begin, end;
  __device__ __forceinline__ uint8_t byte(uint32_t p) const { return p < end ? text[p] : (uint8_t)0; }
  __device__ __forceinline__ bool boundary(uint32_t p) const { return p == begin; }
};
struct ExactSink {
  uint8_t* type;
  uint32_t* start;
  uint32_t* end;
  uint8_t* kw;
  uint32_t cap;
  __device__ __forceinline__ void token(uint32_t i, uint8_t t, uint32_t s, uint32_t e, uint8_t k) {
    if (i < cap) {
      type[i] = t;
      start[i] = s;
      end[i] = e;
      kw[i] = k;
    }
  }
  __device__ __forceinline__ void seg_begin(uint32_t, uint32_t, uint32_t) {}
  __device__ __forceinline__ void seg_end(uint32_t, uint32_t, uint32_t) {}
};

template <bool Emit>
__global__ void __launch_bounds__(128) k_lex_exact(const uint8_t* __restrict__ text, const uint32_t* __restrict__ off32,
                                                   const LexTables* __restrict__ gT, const uint32_t* __restrict__ punt_list,
                                                   const uint32_t* __restrict__ npunt_dev, uint2* __restrict__ counts /* x = tokens, y = 0 */,
                                                   const uint2* __restrict__ offsets, const uint32_t* __restrict__ extra_base_dev,
                                                   ExactSink sink, uint32_t* __restrict__ stmt_tok_begin,
                                                   uint32_t* __restrict__ stmt_tok_end, uint32_t* __restrict__ punt_flag) {
  __shared__ LexTables T;
  const uint32_t npunt = *npunt_dev;  // (the list's length is only known on the device: a fixed grid strides over it)
  if (blockIdx.x * blockDim.x >= npunt) return;
  stage_tables(gT, &T);
  __syncthreads();
  const uint32_t extra_base = Emit ? *extra_base_dev : 0u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npunt; i += gridDim.x * blockDim.x) {
    const uint32_t s = punt_list[i];
    StmtSrc src{text, off32[s], off32[s + 1]};
    LexCarry c;
    c.stmt_start = src.begin;
    c.tok_start = src.begin;
    c.count = Emit ? extra_base + offsets[i].x : 0u;
    const uint32_t first = c.count;
    Walker<false, StmtSrc, ExactSink> w(T, src, sink, c);
    w.counting = !Emit;
    // One thread walks one statement: a byte at a time from global memory would cost a full memory latency per
    // byte.  The statement is read as aligned 16-byte blocks, the next block in flight while this one is walked.
    if (src.begin < src.end) {
      const uint8_t* p0 = text + src.begin;
      const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 15u);
      const uint8_t* blk = p0 - mis;             // (the aligned block of a valid byte lies inside the same allocation)
      const uint8_t* const last = text + src.end - 1u;
      uint4 cur = __ldg(reinterpret_cast<const uint4*>(blk));
      uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
      if (blk + 16 <= last) nxt = __ldg(reinterpret_cast<const uint4*>(blk + 16));
      uint32_t k = mis;
      for (uint32_t pos = src.begin; pos < src.end; pos++, k++) {
        if (k == 16u) {
          k = 0u;
          blk += 16;
          cur = nxt;
          if (blk + 16 <= last) nxt = __ldg(reinterpret_cast<const uint4*>(blk + 16));
        }
        const uint32_t word = k < 8u ? (k < 4u ? cur.x : cur.y) : (k < 12u ? cur.z : cur.w);
        w.step(pos, (uint8_t)((word >> (8u * (k & 3u))) & 255u), pos == src.begin, true);
      }
    }
    w.flush_eof(src.end);
    if (Emit) {
      stmt_tok_begin[s] = first;
      stmt_tok_end[s] = w.c.count;
      punt_flag[s] = i + 1u;  // position in the extra region's order: keeps the parser's node ranges disjoint
    } else {
      counts[i] = make_uint2(w.c.count, 0u);
    }
  }
}
