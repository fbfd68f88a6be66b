// lex2_kernels.cuh -- kernels of the warp-cooperative lexer (logic in lex2_core.cuh).
// Included by nutdb_gpu.cu after the scan helpers and DevSink.
//
//   k_lex2_fn      per warp (1 KB): transition function of the context automaton; per-tile aggregate
//   k_lex2_walk<0> per warp: number of tokens + string / statement carries; flags statements for the exact path
//   k_lex2_walk<1> per warp: emits tokens at their final index
//   k_lex_exact<0/1> one thread per flagged statement: the exact walker of lex_core.cuh (count / emit)
#pragma once

#define L2_THREADS 256
#define L2_WARPS (L2_THREADS / 32)
#define L2_SEG 1024                     // bytes per warp
#define L2_TILE (L2_WARPS * L2_SEG)     // bytes per block

struct alignas(16) Lex2Shared {
  alignas(16) uint32_t text[L2_TILE / 4];  // filled with 16-byte stores
  uint32_t bm[L2_TILE / 32];
  LexTables T;
  nlex2::Lex2Tables K;
};

struct Tile2Src {
  const uint8_t* text;
  const uint8_t* sm;
  uint32_t tile_begin, n;
  __device__ __forceinline__ uint8_t byte(uint32_t p) const {
    const uint32_t r = p - tile_begin;
    if (r < L2_TILE) return sm[r];
    return p < n ? text[p] : (uint8_t)0;
  }
};

__device__ __forceinline__ void stage_tile2(const uint8_t* text, const uint32_t* bitmap, uint32_t tile_begin, uint32_t n,
                                            Lex2Shared& S, const LexTables* gT, const nlex2::Lex2Tables* gK) {
  stage_tables(gT, &S.T);
  {
    const uint32_t* a = reinterpret_cast<const uint32_t*>(gK);
    uint32_t* b = reinterpret_cast<uint32_t*>(&S.K);
    for (uint32_t i = threadIdx.x; i < sizeof(nlex2::Lex2Tables) / 4; i += blockDim.x) b[i] = a[i];
  }
  const uint4* src = reinterpret_cast<const uint4*>(text + tile_begin);
  uint4* dst = reinterpret_cast<uint4*>(S.text);
#pragma unroll
  for (uint32_t k = threadIdx.x; k < L2_TILE / 16; k += L2_THREADS) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (tile_begin + 16u * k < n) v = __ldg(src + k);
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

// stage 1: raw class masks of the window at `base` (warp-wide)
__device__ __forceinline__ void build_win(const Lex2Shared& S, const Tile2Src& src, const uint32_t* bitmap,
                                          uint32_t tile_begin, uint32_t base, uint32_t n, int lane, uint8_t& b, uint16_t& k,
                                          nlex2::Win& w) {
  const uint32_t pos = base + (uint32_t)lane;
  b = src.byte(pos);
  k = pos < n ? S.K.cls[b] : (uint16_t)0;
  const uint32_t full = 0xFFFFFFFFu;
  w.valid = base + 32u <= n ? full : (n > base ? ((1u << (n - base)) - 1u) : 0u);
  w.sq = __ballot_sync(full, k & nlex2::K_SQ);
  w.dq = __ballot_sync(full, k & nlex2::K_DQ);
  w.bt = __ballot_sync(full, k & nlex2::K_BT);
  w.nl = __ballot_sync(full, k & nlex2::K_NL);
  w.bs = __ballot_sync(full, k & nlex2::K_BS);
  w.dash = __ballot_sync(full, k & nlex2::K_DASH);
  w.slash = __ballot_sync(full, k & nlex2::K_SLASH);
  w.star = __ballot_sync(full, k & nlex2::K_STAR);
  w.L = __ballot_sync(full, k & nlex2::K_L);
  w.D = __ballot_sync(full, k & nlex2::K_D);
  w.DOT = __ballot_sync(full, k & nlex2::K_DOT);
  w.OP = __ballot_sync(full, k & nlex2::K_OP);
  uint32_t bnd = bnd_word(S, bitmap, tile_begin, base) & w.valid;
  if (n >= base && n - base < 32u) bnd |= 1u << (n - base);  // the batch end terminates the last statement
  w.bnd = bnd;
}

__device__ __forceinline__ nlex2::Next next_of(const Lex2Shared& S, const Tile2Src& src, const uint32_t* bitmap,
                                               uint32_t tile_begin, uint32_t base, uint32_t n) {
  nlex2::Next nx;
  const uint32_t p = base + 32u;
  if (p >= n) {
    nx.byte = 0;
    nx.bnd = 1;
    nx.cls = 0;
  } else {
    nx.byte = src.byte(p);
    nx.bnd = (uint8_t)(bnd_word(S, bitmap, tile_begin, p) & 1u);
    nx.cls = S.K.cls[nx.byte];
  }
  return nx;
}

// backslash parity and previous byte in front of a segment
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
    if ((bitmap[p >> 5] >> (p & 31u)) & 1u) break;  // the run cannot extend over a statement start
  }
  esc = (uint8_t)(nrun & 1u);
}

__global__ void __launch_bounds__(L2_THREADS) k_lex2_fn(const uint8_t* __restrict__ text, const uint32_t* __restrict__ bitmap,
                                                        uint32_t n, const LexTables* __restrict__ gT,
                                                        const nlex2::Lex2Tables* __restrict__ gK,
                                                        uint32_t* __restrict__ localA, uint32_t* __restrict__ tileA) {
  __shared__ Lex2Shared S;
  __shared__ uint32_t wfn[L2_WARPS];
  const uint32_t tile_begin = blockIdx.x * L2_TILE;
  stage_tile2(text, bitmap, tile_begin, n, S, gT, gK);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Tile2Src src{text, reinterpret_cast<const uint8_t*>(S.text), tile_begin, n};
  const uint32_t seg = tile_begin + (uint32_t)warp * L2_SEG;
  uint32_t run = NUTDB_VEC8_ID;
  if (seg < n) {
    uint8_t prev, esc;
    entry_esc(src, bitmap, seg, prev, esc);
    const uint32_t end = min(seg + L2_SEG, n);
    for (uint32_t base = seg; base < end; base += 32u) {
      const uint32_t pos = base + (uint32_t)lane;
      const uint8_t b = src.byte(pos);
      const uint16_t k = pos < n ? S.K.cls[b] : (uint16_t)0;
      const uint32_t full = 0xFFFFFFFFu;
      nlex2::Win w;
      w.valid = base + 32u <= n ? full : ((1u << (n - base)) - 1u);
      w.sq = __ballot_sync(full, k & nlex2::K_SQ);
      w.dq = __ballot_sync(full, k & nlex2::K_DQ);
      w.bt = __ballot_sync(full, k & nlex2::K_BT);
      w.nl = __ballot_sync(full, k & nlex2::K_NL);
      w.bs = __ballot_sync(full, k & nlex2::K_BS);
      w.dash = __ballot_sync(full, k & nlex2::K_DASH);
      w.slash = __ballot_sync(full, k & nlex2::K_SLASH);
      w.star = __ballot_sync(full, k & nlex2::K_STAR);
      w.L = w.D = w.DOT = w.OP = 0;
      w.bnd = bnd_word(S, bitmap, tile_begin, base) & w.valid;
      const uint32_t escm = __ballot_sync(full, nlex2::lane_esc(w.bs, lane, esc)) & ~w.bnd;
      const nlex2::Events ev = nlex2::make_events(w, escm, prev);
      if (ev.all) run = nlex2::ctx_window_fn(S.T, w, ev, run);
      else run = vec8_then_row(run, S.T.a_row[EV_OTHER][0], S.T.a_row[EV_OTHER][1]);
      esc = nlex2::esc_carry_out(w.bs, esc);
      prev = (uint8_t)__shfl_sync(full, (uint32_t)b, 31);
    }
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
  uint32_t* punt_list;   // statement indices
  uint32_t* punt_count;
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
    if (atomicExch(&punt_flag[s], 1u) == 0u) punt_list[atomicAdd(punt_count, 1u)] = s;
  }
};

template <bool Emit>
__global__ void __launch_bounds__(L2_THREADS) k_lex2_walk(const uint8_t* __restrict__ text,
                                                          const uint32_t* __restrict__ bitmap, uint32_t n,
                                                          const LexTables* __restrict__ gT,
                                                          const nlex2::Lex2Tables* __restrict__ gK,
                                                          const uint32_t* __restrict__ localA,
                                                          const uint8_t* __restrict__ tileEntA, uint4* __restrict__ localC,
                                                          uint4* __restrict__ tileC, const uint4* __restrict__ tilePrefC,
                                                          Lex2Out out) {
  __shared__ Lex2Shared S;
  __shared__ uint4 wsum[L2_WARPS];
  const uint32_t tile_begin = blockIdx.x * L2_TILE;
  stage_tile2(text, bitmap, tile_begin, n, S, gT, gK);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t full = 0xFFFFFFFFu, lt = (1u << lane) - 1u;
  Tile2Src src{text, reinterpret_cast<const uint8_t*>(S.text), tile_begin, n};
  const uint32_t seg = tile_begin + (uint32_t)warp * L2_SEG;
  const uint32_t widx = blockIdx.x * L2_WARPS + warp;
  nlex2::Carry2 c;
  bool opened = false;
  uint32_t nbnd_seen = 0;
  if (seg < n) {
    c.s = (uint8_t)vec8_apply(localA[widx], tileEntA[blockIdx.x]);
    entry_esc(src, bitmap, seg, c.prev, c.esc);
    c.str_start = NUTDB_NO_TOK;
    c.stmt_start = 0;
    if (Emit) {
      const CSum pre = csum_unpack(CSumOp::then(tilePrefC[blockIdx.x], localC[widx]));
      c.count = pre.count;
      c.str_start = pre.tok_start;
      c.escaped = pre.escaped;
      c.stmt_start = pre.stmt_start;
    }
    // history: class masks of the 32 bytes in front of the segment (raw masks are exact there when we enter in code)
    nlex2::Hist h;
    if (seg >= 32u && c.s <= A_CX) {
      const uint32_t pos = seg - 32u + (uint32_t)lane;
      const uint16_t k = S.K.cls[src.byte(pos)];
      h.L = __ballot_sync(full, k & nlex2::K_L);
      h.D = __ballot_sync(full, k & nlex2::K_D);
      h.DOT = __ballot_sync(full, k & nlex2::K_DOT);
      h.OP = __ballot_sync(full, k & nlex2::K_OP);
      h.bnd = bitmap[(seg - 32u) >> 5];
    } else if (seg >= 32u) {
      h.bnd = bitmap[(seg - 32u) >> 5];
    }
    {  // '' / "" split exactly at the segment start
      const uint8_t b0 = src.byte(seg);
      const bool bnd0 = (bnd_word(S, bitmap, tile_begin, seg) & 1u) != 0;
      c.reopen = (c.s == A_C && !bnd0 && (c.prev == '\'' || c.prev == '"') && b0 == c.prev) ? 1 : 0;
    }
    const uint32_t end = min(seg + L2_SEG, n);
    for (uint32_t base = seg; base < end; base += 32u) {
      uint8_t b;
      uint16_t k;
      nlex2::Win w;
      build_win(S, src, bitmap, tile_begin, base, n, lane, b, k, w);
      const nlex2::Next nx = next_of(S, src, bitmap, tile_begin, base, n);
      const uint32_t escm = __ballot_sync(full, nlex2::lane_esc(w.bs, lane, c.esc)) & ~w.bnd;
      const nlex2::Events ev = nlex2::make_events(w, escm, c.prev);
      const uint32_t stmt_entry = c.stmt_start;
      const uint32_t str_before = c.str_start;
      nlex2::CtxOut o;
      nlex2::ctx_window(w, ev, base, nx, lane, c, o);
      if (c.str_start != str_before) opened = true;
      nbnd_seen += (uint32_t)__popc(w.bnd & w.valid);
      nlex2::LaneTok t = nlex2::lane_token(S.T, src, lane, base, b, k, w, o, h, nx, escm, c.prev);
      const uint32_t tokmask = __ballot_sync(full, t.has);
      const uint32_t eofmask = __ballot_sync(full, t.eof);
      if (!Emit) {
        const uint32_t badmask = __ballot_sync(full, t.bad) | o.bad;
        if (badmask | o.bad_prev) {
          if ((badmask >> lane) & 1u) out.punt(base + (uint32_t)lane);
          if ((o.bad_prev >> lane) & 1u) out.punt(base + (uint32_t)lane - 1u);
        }
      } else {
        const uint32_t idx = c.count + (uint32_t)__popc(tokmask & lt) + (uint32_t)__popc(eofmask & lt);
        // start of the statement this lane's token belongs to
        const uint32_t below = w.bnd & w.valid & (lt | (1u << lane));
        const uint32_t sst = below ? base + (uint32_t)(31 - __clz((int)below)) : stmt_entry;
        if ((w.bnd & w.valid) & (1u << lane)) out.stmt_tok_begin[out.find_stmt(base + (uint32_t)lane)] = idx;
        if (t.has && idx < out.cap) {
          out.type[idx] = t.type;
          out.start[idx] = t.start - sst;
          out.end[idx] = t.end - sst;
          out.kw[idx] = t.kw;
        }
        if (t.eof) {
          const uint32_t e = idx + (uint32_t)t.has;
          if (e < out.cap) {
            out.type[e] = NUTDB_TT_EOF;
            out.start[e] = base + (uint32_t)lane + 1u - sst;
            out.end[e] = base + (uint32_t)lane + 1u - sst;
            out.kw[e] = 0;
          }
          out.stmt_tok_end[out.find_stmt(base + (uint32_t)lane)] = e + 1u;
        }
      }
      c.count += (uint32_t)__popc(tokmask) + (uint32_t)__popc(eofmask);
      // next window's history and carries
      h.L = w.L & o.ct;
      h.D = w.D & o.ct;
      h.DOT = w.DOT & o.ct;
      h.OP = w.OP & o.ct;
      h.bnd = w.bnd;
      c.esc = nlex2::esc_carry_out(w.bs, c.esc);
      c.prev = (uint8_t)__shfl_sync(full, (uint32_t)b, 31);
    }
    // the batch ends exactly on a window boundary: no window carries the virtual end-of-batch statement
    // start, so check here that the last statement did not end inside a string / quoted identifier / comment
    if (!Emit && end == n && (n & 31u) == 0u && lane == 0 &&
        (c.s == A_SQ || c.s == A_DQ || c.s == A_BT || c.s == A_BC0 || c.s == A_BC))
      out.punt(n - 1u);
  }
  if (!Emit) {
    if (lane == 0) {
      CSum s;
      s.count = c.count;
      s.nseg = nbnd_seen;
      s.has_tok = opened ? 1 : 0;
      s.tok_start = c.str_start;
      s.escaped = c.escaped;
      s.stmt_start = c.stmt_start;
      wsum[warp] = csum_pack(s);
    }
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
                                                   uint32_t npunt, uint2* __restrict__ counts /* x = tokens, y = 0 */,
                                                   const uint2* __restrict__ offsets, uint32_t extra_base, ExactSink sink,
                                                   uint32_t* __restrict__ stmt_tok_begin, uint32_t* __restrict__ stmt_tok_end,
                                                   uint32_t* __restrict__ punt_flag) {
  __shared__ LexTables T;
  stage_tables(gT, &T);
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npunt) return;
  const uint32_t s = punt_list[i];
  StmtSrc src{text, off32[s], off32[s + 1]};
  LexCarry c;
  c.stmt_start = src.begin;
  c.tok_start = src.begin;
  c.count = Emit ? extra_base + offsets[i].x : 0u;
  const uint32_t first = c.count;
  Walker<false, StmtSrc, ExactSink> w(T, src, sink, c);
  w.counting = !Emit;
  for (uint32_t pos = src.begin; pos < src.end; pos++) w.step(pos, src.byte(pos), pos == src.begin, true);
  w.flush_eof(src.end);
  if (Emit) {
    stmt_tok_begin[s] = first;
    stmt_tok_end[s] = w.c.count;
    punt_flag[s] = i + 1u;  // position in the extra region's order: keeps the parser's node ranges disjoint
  } else {
    counts[i] = make_uint2(w.c.count, 0u);
  }
}
