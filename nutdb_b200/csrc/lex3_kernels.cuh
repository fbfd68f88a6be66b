// lex3_kernels.cuh -- k_lex3: the whole lexer in ONE pass over the text (logic: lex3_core.cuh).
//
// Stages 1-3 of the pipeline in a single persistent kernel: the text is read once, nothing but the token arrays
// and 12 bytes per 32-byte window (token index + token-end masks, from which the parser derives every statement's
// token range) is written.
//
//   * persistent CTAs (one 8 KB tile per iteration, tiles handed out by an atomic ticket, so a tile's predecessors
//     are always running or finished -- the forward-progress guarantee the look-back needs);
//   * the next tile is staged while this one is lexed: one thread issues a bulk asynchronous copy
//     (cp.async.bulk global -> shared, completion counted on an mbarrier), two buffers;
//   * two DECOUPLED LOOK-BACK scans chained inside the kernel: (a) the context automaton's transition function
//     (8 nibbles; ends early as soon as the composed function is constant, which a statement start makes it),
//     (b) token count + open-literal carry + statement start;
//   * thread per WINDOW for masks / context walk / token records, then thread per TOKEN for type, keyword hash,
//     end-of-token checks and coalesced stores (records staged in shared memory).
#pragma once
#include "lex3_core.cuh"

#ifndef L3_WORKERS
#define L3_WORKERS 256                    // worker threads: one 32-byte window each
#endif
#define L3_THREADS (L3_WORKERS + 32)      // + one helper warp that runs the two look-back scans beside them
#define L3_WARPS (L3_WORKERS / 32)
#define L3_WIN L3_WORKERS                 // windows per tile
#define L3_TILE (L3_WIN * 32)             // bytes per tile
#define L3_HALO 32
#ifndef L3_RCAP
#define L3_RCAP 2560                      // token records staged per round (a tile holds ~2000 on query logs)
#endif
#ifndef L3_MINBLOCKS
#define L3_MINBLOCKS 3
#endif
#ifndef L3_LB_GROUPS
#define L3_LB_GROUPS 2                    // look-back (b): 32 x this many predecessors per polling step
#endif

struct alignas(16) Lex3Shared {
  alignas(16) uint8_t pad[16];             // (aligned word reads around a token may begin 3 bytes in front of text[0])
  alignas(16) uint8_t text[2][L3_HALO + L3_TILE + L3_HALO];
  alignas(16) uint32_t bm[2][L3_WIN + 4];  // statement-start bitmap words of the tile and of the window after it
  alignas(8) unsigned long long bar[2];
  LexTables T;
  uint32_t bndm[L3_WIN + 1];               // statement starts per window incl. the virtual one at the batch end
  uint32_t sst_in[L3_WIN];                 // start of the statement that is open where the window begins
  uint32_t rec0[L3_RCAP], rec1[L3_RCAP];
  uint32_t wfn[L3_WARPS];
  uint4 wsum[L3_WARPS];
  uint32_t ticket[2];
  uint32_t s_tile_in;                      // look-back (a): context state at the tile's first byte
  uint4 c_tile_pre;                        // look-back (b): tokens / statement start / open literal before the tile
  volatile uint32_t agg_ready;             // tile + 1 once wsum holds this tile's sums
  uint32_t twin_entry;                     // the tile begins with the second quote of '' / "" (the literal goes on)
};

// Look-back descriptors, one 64-bit word per tile each.  Every word validates itself: its high half is a tag (launch
// epoch, status 1 = the tile's own contribution, 2 = inclusive of all tiles before it), so no fences are needed and
// nothing has to be cleared between launches.
//   fn : tag = epoch << 2 | status; low = context transition function (8 nibbles).  A statement start in the tile
//        makes its own function constant, which already is the inclusive state: most walks are one hop.
//   cA : tag = epoch << 2 | 1; low = tokens of the tile.  cI : tag = epoch << 2 | 2; low = tokens up to and including
//        the tile -- written by the SCANNER (block 0), see l3_scanner.
//   yd : tag = epoch << 2 | status; low = 1 + last statement start (status 2), or nothing (status 1: the tile has no
//        statement start -- look further back)
//   zd : tag = epoch << 5 | status << 3 | flags; low = offset of the last opening quote.  Status 2: flags = 1 |
//        escaped << 1 | backslash-u << 2; status 1 (no literal opened in the tile): flags = its contribution to the
//        escaped and backslash-u flags
struct Lex3Desc {
  unsigned long long* fn;
  unsigned long long* cA;
  unsigned long long* cI;
  unsigned long long* yd;
  unsigned long long* zd;
};

struct Lex3Out {
  uint8_t* type;
  uint32_t* start;
  uint32_t* end;
  uint8_t* kw;
  uint32_t cap;            // capacity of the token arrays (tokens beyond it are counted, not stored)
  uint32_t* win_idx;       // per window: index of the first token that ends in it (+1 entry: the total)
  uint32_t* win_has;       // per window: bytes where a token ends
  uint32_t* win_eof;       // per window: bytes after which an EOF token follows
  const uint32_t* off32;
  uint32_t nstmt;
  uint32_t* punt_flag;     // per statement: needs the exact lexer
  uint32_t* counters;      // [0] flagged statements, [1] bound on their tokens, [2] total tokens, [3] tile ticket
  const uint32_t* first_stmt;
  uint32_t* dbg;           // optional (tests): per tile {entry state, first token index, statement start, open quote}
  unsigned long long* tim; // optional (profiling): 24 clock stamps per tile (12 worker warp 0, 12 helper)
  __device__ void punt_stmt_at(uint32_t sst) const {  // the (non-empty) statement that starts at byte sst
    uint32_t c = first_stmt[sst >> 5];
    while (off32[c] != sst || off32[c + 1] == sst) c++;
    if (atomicExch(&punt_flag[c], 1u) == 0u) {
      atomicAdd(counters, 1u);
      atomicAdd(counters + 1, off32[c + 1] - off32[c] + 1u);
    }
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One thread stages tile t (text with a 32-byte halo on both sides + its bitmap words) into buffer b.
// `padded` = bytes readable behind the text (the library's own buffer is padded; a caller's device buffer is not).
__device__ __forceinline__ void l3_issue_load(Lex3Shared& S, int b, uint32_t t, const uint8_t* text, const uint32_t* bitmap,
                                              uint32_t n, uint32_t n_readable) {
  const uint32_t tile_begin = t * L3_TILE;
  const uint32_t lo = tile_begin >= L3_HALO ? tile_begin - L3_HALO : 0u;
  uint32_t hi = tile_begin + L3_TILE + L3_HALO;
  if (hi > n_readable) hi = n_readable & ~15u;  // whole 16-byte pieces only; the ragged rest is loaded by threads
  const uint32_t bytes_text = hi > lo ? hi - lo : 0u;
  const uint32_t bytes_bm = (L3_WIN + 4) * 4u;    // (the bitmap has that many words behind every tile)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(&S.bar[b], bytes_text + bytes_bm);
  if (bytes_text) bulk_g2s(&S.text[b][L3_HALO + lo - tile_begin], text + lo, bytes_text, &S.bar[b]);
  bulk_g2s(&S.bm[b][0], bitmap + (tile_begin >> 5), bytes_bm, &S.bar[b]);
}

struct Tile3Src {
  const uint8_t* text;
  const uint8_t* sm;  // the tile's first byte in shared memory; the halo lies at sm[-32..-1] and sm[L3_TILE..+31]
  uint32_t tile_begin, n;
  __device__ __forceinline__ uint8_t byte(uint32_t p) const {
    const uint32_t r = p - tile_begin + L3_HALO;
    if (r < L3_TILE + 2 * L3_HALO) return sm[(int)r - L3_HALO];
    return far_byte(text, p, n);
  }
  // bytes of a code token that ends in this tile: always staged (it starts at most 32 bytes in front of the tile)
  __device__ __forceinline__ uint8_t at(uint32_t p) const { return sm[(int)(p - tile_begin)]; }
  __device__ __forceinline__ const uint8_t* span(uint32_t p, uint32_t) const { return sm + (int)(p - tile_begin); }
};

__device__ __forceinline__ uint4 c3_then(const uint4& a, const uint4& b) {
  uint4 r;
  r.x = a.x + b.x;
  r.y = b.y ? b.y : a.y;
  if (b.w & 1u) {
    r.z = b.z;
    r.w = b.w;
  } else {
    r.z = a.z;
    r.w = a.w | (b.w & (2u | 8u));  // escaped flag, backslash-u flag
  }
  return r;
}
__device__ __forceinline__ uint4 c3_shfl_up(const uint4& v, int d) {
  return make_uint4(__shfl_up_sync(0xFFFFFFFFu, v.x, d), __shfl_up_sync(0xFFFFFFFFu, v.y, d),
                    __shfl_up_sync(0xFFFFFFFFu, v.z, d), __shfl_up_sync(0xFFFFFFFFu, v.w, d));
}
__device__ __forceinline__ bool vec8_is_const(uint32_t f) { return f == (f & 7u) * 0x11111111u; }
// A load the compiler may not hoist out of (or delete together with) a polling loop: a plain / __ldcg load is
// loop-invariant to it, and a loop without side effects "must terminate", so the epoch test would be removed.
__device__ __forceinline__ unsigned long long ld_poll_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_pub_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 0, %0;" ::"n"(L3_THREADS) : "memory"); }
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(L3_WORKERS) : "memory"); }

// look-back (a): the context state at the first byte of `tile` = the functions of all tiles before it applied to A_C
__device__ __forceinline__ uint32_t l3_lookback_state(const Lex3Desc& desc, uint32_t epoch, uint32_t tile) {
  if (tile == 0) return A_C;
  uint32_t acc = NUTDB_VEC8_ID;  // functions of the tiles (p, tile) composed
  uint32_t p = tile;
  for (;;) {
    p--;
    unsigned long long d;
    for (;;) {
      d = ld_poll_u64(desc.fn + p);
      if ((uint32_t)(d >> 34) == epoch) break;
      __nanosleep(100);  // (a polling warp must not take issue slots from the warps that produce what it waits for)
    }
    acc = vec8_then((uint32_t)d, acc);
    if (((d >> 32) & 3ull) == 2ull || vec8_is_const(acc) || p == 0) break;
  }
  return vec8_apply(acc, A_C);
}
// "last statement start" / "open quote" before `tile`: walks back over the tiles that have none (usually one hop)
__device__ __forceinline__ uint32_t l3_lookback_y(const Lex3Desc& desc, uint32_t epoch, uint32_t tile) {
  uint32_t p = tile;
  while (p > 0) {
    p--;
    unsigned long long d;
    for (;;) {
      d = ld_poll_u64(desc.yd + p);
      if ((uint32_t)(d >> 34) == epoch && ((d >> 32) & 3ull) != 0ull) break;
      __nanosleep(100);
    }
    if (((d >> 32) & 3ull) == 2ull) return (uint32_t)d;
  }
  return 0u;
}
// carry word (bit 0 open, 1 escaped, 3 backslash-u) <-> the three descriptor flag bits
__device__ __forceinline__ uint32_t l3_zflags(uint32_t w) { return (w & 3u) | ((w & 8u) >> 1); }
__device__ __forceinline__ uint32_t l3_zcarry(uint32_t f) { return (f & 3u) | ((f & 4u) << 1); }
__device__ __forceinline__ void l3_lookback_z(const Lex3Desc& desc, uint32_t epoch, uint32_t tile, uint32_t& z, uint32_t& w) {
  uint32_t p = tile, esc = 0;
  z = 0;
  w = 0;
  while (p > 0) {
    p--;
    unsigned long long d;
    uint32_t tag;
    for (;;) {
      d = ld_poll_u64(desc.zd + p);
      tag = (uint32_t)(d >> 32);
      if ((tag >> 5) == epoch && ((tag >> 3) & 3u) != 0u) break;
      __nanosleep(100);
    }
    if (((tag >> 3) & 3u) == 2u) {
      z = (uint32_t)d;
      w = l3_zcarry(tag & 7u) | esc;
      return;
    }
    esc |= l3_zcarry(tag & 6u);
  }
  w = esc;
}
__device__ __forceinline__ unsigned long long l3_tag2(uint32_t epoch, uint32_t status) {
  return (unsigned long long)((epoch << 2) | status) << 32;
}

// The token-count scan.  Letting every tile look back over its predecessors' counts makes hundreds of CTAs poll the
// same few descriptor lines (measured: the L2 slices holding them serialise the polls, ~5 us per look-back step, and
// the whole kernel runs at the pace of that chain).  Instead ONE warp (block 0) turns the tiles' counts into inclusive
// prefixes as they arrive -- up to 128 tiles per step, each count read by nobody else -- and a tile only polls the
// single word its predecessor's prefix lands in.
__device__ __forceinline__ void l3_scanner(const Lex3Desc& desc, uint32_t epoch, uint32_t ntiles, const Lex3Out& out) {
  if (threadIdx.x >= 32u) return;
  const uint32_t lane = threadIdx.x, full = 0xFFFFFFFFu;
  uint32_t base = 0, running = 0;
  while (base < ntiles) {
    uint32_t v[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t i = base + lane + 32u * j;
      v[j] = 0;
      ok[j] = true;  // (beyond the last tile: nothing to wait for)
      if (i < ntiles) {
        const unsigned long long a = ld_poll_u64(desc.cA + i);
        ok[j] = (uint32_t)(a >> 34) == epoch && ((a >> 32) & 3ull) != 0ull;
        if (ok[j]) v[j] = (uint32_t)a;
      }
    }
    uint32_t f = 128u;  // the first tile of the window whose count has not arrived
#pragma unroll
    for (int j = 3; j >= 0; j--) {
      const uint32_t m = __ballot_sync(full, !ok[j]);
      if (m) f = 32u * j + (uint32_t)(__ffs((int)m) - 1);
    }
    f = min(f, ntiles - base);
    if (f == 0u) {
      __nanosleep(100);
      continue;
    }
    uint32_t carry = running;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const bool in = 32u * j + lane < f;
      uint32_t x = in ? v[j] : 0u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(full, x, d);
        if (lane >= (uint32_t)d) x += y;
      }
      if (in) st_pub_u64(desc.cI + base + 32u * j + lane, ((unsigned long long)((epoch << 2) | 2u) << 32) | (carry + x));
      carry += __shfl_sync(full, x, 31);
    }
    running = carry;
    base += f;
  }
  if (lane == 0) {
    out.counters[2] = running;
    out.win_idx[(size_t)ntiles * L3_WIN] = running;
  }
}

__global__ void __launch_bounds__(L3_THREADS, L3_MINBLOCKS) k_lex3(const uint8_t* __restrict__ text,
                                                                    const uint32_t* __restrict__ bitmap, uint32_t n,
                                                                    uint32_t n_readable, uint32_t ntiles,
                                                                    const LexTables* __restrict__ gT, Lex3Desc desc,
                                                                    uint32_t epoch, Lex3Out out,
                                                                    const uint32_t* __restrict__ gate) {
  if (*gate) return;  // invalid statement offsets (k_prep): nothing downstream may trust them
  if (blockIdx.x == 0) {
    l3_scanner(desc, epoch, ntiles, out);
    return;
  }
  extern __shared__ __align__(16) unsigned char l3_smem[];
  Lex3Shared& S = *reinterpret_cast<Lex3Shared*>(l3_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool helper = warp == L3_WARPS;
  const uint32_t full = 0xFFFFFFFFu;
  stage_tables(gT, &S.T);
  if (threadIdx.x == L3_WORKERS) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    S.agg_ready = 0u;
    const uint32_t t0 = atomicAdd(out.counters + 3, 1u);
    S.ticket[0] = t0;
    if (t0 < ntiles) l3_issue_load(S, 0, t0, text, bitmap, n, n_readable);
  }
  __syncthreads();
  uint32_t phase0 = 0u, phase1 = 0u;
  for (int b = 0;; b ^= 1) {
    const uint32_t tile = S.ticket[b];
    if (tile >= ntiles) break;
    const uint32_t tile_begin = tile * L3_TILE;
    uint8_t* const sm = &S.text[b][L3_HALO];
    const uint32_t* const bm = S.bm[b];
    Tile3Src src{text, sm, tile_begin, n};
    // per-thread state that lives across the barriers (workers)
    const uint32_t base = tile_begin + 32u * threadIdx.x;
    const bool live = !helper && base < n;
    nlex3::WinCtx3 o;
    nlex3::TokMasks m;
    uint4 mine = make_uint4(0u, 0u, 0u, 0u), excl = make_uint4(0u, 0u, 0u, 0u);
    uint32_t local = 0;
    bool deferred = false;

#define L3_STAMP(slot) do { if (out.tim && lane == 0 && (warp == 0 || helper)) out.tim[24 * (size_t)tile + (helper ? 12 : 0) + (slot)] = clock64(); } while (0)
    L3_STAMP(0);
    if (helper) {
      // ======== helper warp: staging of the next tile, both look-back scans ========
      if (lane == 0) {  // claim and stage the next tile while this one is lexed
        const uint32_t tn = atomicAdd(out.counters + 3, 1u);
        S.ticket[b ^ 1] = tn;
        if (tn < ntiles) l3_issue_load(S, b ^ 1, tn, text, bitmap, n, n_readable);
      }
      uint32_t s_in = 0;
      if (lane == 0) {  // (a) depends on the tiles before this one only: it runs beside the workers' stage 1
#ifdef L3_NOWAIT
        s_in = A_C;
#else
        s_in = l3_lookback_state(desc, epoch, tile);
#endif
        S.s_tile_in = s_in;
      }
      L3_STAMP(1);
      bar_all();  // ---- A: the windows' transition functions are in wfn (a worker has published the tile's own function)
      if (lane == 0) {
        uint32_t agg = NUTDB_VEC8_ID;
#pragma unroll
        for (int i = 0; i < L3_WARPS; i++) agg = vec8_then(agg, S.wfn[i]);
        if (!vec8_is_const(agg))
          st_pub_u64(desc.fn + tile, ((unsigned long long)((epoch << 2) | 2u) << 32) | (vec8_apply(agg, s_in) * 0x11111111u));
      }
      // (b): tokens before the tile = the scanner's inclusive prefix of the tile before this one
      uint32_t pre_count = 0;
      bool published = false;
      auto publish_own = [&](bool) {  // the tile's own sums (needs agg_ready)
        __threadfence_block();
        uint4 agg = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < L3_WARPS; i++) agg = c3_then(agg, S.wsum[i]);
        if (lane == 0 && !published) {
          st_pub_u64(desc.cA + tile, l3_tag2(epoch, 1u) | (unsigned long long)agg.x);
          st_pub_u64(desc.yd + tile, l3_tag2(epoch, agg.y ? 2u : 1u) | agg.y);
          st_pub_u64(desc.zd + tile, ((unsigned long long)((epoch << 5) | ((agg.w & 1u) ? 16u : 8u) | l3_zflags(agg.w)) << 32) | agg.z);
        }
        published = true;
        return agg;
      };
      L3_STAMP(2);
      while (S.agg_ready != tile + 1u) __nanosleep(40);
      L3_STAMP(3);
      {
        const uint4 agg = publish_own(true);
        if (tile > 0) {
          if (lane == 0) {
            unsigned long long a;
            for (;;) {
              a = ld_poll_u64(desc.cI + tile - 1u);
              if ((uint32_t)(a >> 34) == epoch) break;
              __nanosleep(60);
            }
            pre_count = (uint32_t)a;
          }
          pre_count = __shfl_sync(full, pre_count, 0);
        }
        // Statement start / open literal before the tile: lanes 0 and 1 walk back over the tiles that have none
        // (the open literal only matters -- and is only looked for -- when the tile begins inside one).
        uint32_t y_in = 0, z_in = 0, w_in = 0;
        const uint32_t s_tile = __shfl_sync(full, s_in, 0);
        // ("inside" includes a tile that begins with the second quote of a doubled one: the literal continues)
        const bool in_literal = s_tile == A_SQ || s_tile == A_DQ || s_tile == A_BT || (s_tile == A_C && S.twin_entry);
#ifndef L3_NOWAIT
        if (lane == 0) y_in = l3_lookback_y(desc, epoch, tile);
        if (lane == 1 && in_literal) l3_lookback_z(desc, epoch, tile, z_in, w_in);
#endif
        z_in = __shfl_sync(full, z_in, 1);
        w_in = __shfl_sync(full, w_in, 1);
        if (lane == 0) {
          // a tile without a statement start / an opening quote passes on what it found: later walks end here
          if (!agg.y) st_pub_u64(desc.yd + tile, l3_tag2(epoch, 2u) | y_in);
          if (in_literal && !(agg.w & 1u))
            st_pub_u64(desc.zd + tile, ((unsigned long long)((epoch << 5) | 16u | l3_zflags(w_in | agg.w)) << 32) | z_in);
          S.c_tile_pre = make_uint4(pre_count, y_in, z_in, w_in);
        }
      }
      // (the text is needed for the token stage only; waiting on the barrier also makes the bulk copy's bytes visible)
      mbar_wait(&S.bar[b], b ? phase1 : phase0);
    } else {
      // ======== workers ========
      mbar_wait(&S.bar[b], b ? phase1 : phase0);
      // what the bulk copy could not bring: the front halo of tile 0, and everything behind the last whole 16-byte
      // piece of a caller's unpadded buffer (loaded bytewise up to n, zero beyond)
      if (tile_begin == 0 && threadIdx.x < L3_HALO) sm[(int)threadIdx.x - L3_HALO] = 0;
      if (tile_begin + L3_TILE + L3_HALO > n_readable) {
        const uint32_t from = n_readable & ~15u;
        for (uint32_t p = max(from, tile_begin >= L3_HALO ? tile_begin - L3_HALO : 0u) + threadIdx.x;
             p < tile_begin + L3_TILE + L3_HALO; p += L3_WORKERS)
          sm[(int)(p - tile_begin)] = p < n ? text[p] : (uint8_t)0;
        bar_workers();
      } else if (tile_begin == 0) {
        bar_workers();
      }
      L3_STAMP(1);
      const uint32_t blk = tile_begin + 1024u * (uint32_t)warp;
      // ---------------- stage 1: class masks, escapes, context events, transition function ----------------
      nlex2::Win w;
      nlex3::Ops op;
      w.valid = base + 32u <= n ? full : (n > base ? ((1u << (n - base)) - 1u) : 0u);
      {
        const uint4* wp = reinterpret_cast<const uint4*>(sm + 32u * threadIdx.x);
        const uint4 q0 = wp[0], q1 = wp[1];
        const uint32_t v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t p[8];
        nlex3::bit_planes(v, p);
        nlex3::classify_planes(p, w.valid, w, op);
      }
      {
        uint32_t bnd = base < n ? (bm[threadIdx.x] & w.valid) : 0u;
        if (n >= base && n - base < 32u) bnd |= 1u << (n - base);  // the batch end terminates the last statement
        w.bnd = bnd;
      }
      nlex2::Next nx;
      {
        const uint32_t p = base + 32u;
        if (p >= n) {
          nx.byte = 0;
          nx.bnd = 1;
          nx.cls = 0;
        } else {
          nx.byte = sm[32u * threadIdx.x + 32u];
          nx.bnd = (uint8_t)(bm[threadIdx.x + 1] & 1u);
          const uint8_t pr = S.T.prop[nx.byte];
          nx.cls = (uint16_t)(((pr & PR_IDENT_END) ? 1u : 0u) | ((pr & PR_NUM_END) ? 2u : 0u));
        }
      }
      uint8_t prev_byte, prev2_byte, esc_in;
      {
        const uint32_t bs_prev = __shfl_up_sync(full, w.bs, 1);
        if (lane == 0) {
          prev_byte = 0;
          esc_in = 0;
          if (base > 0 && base <= n) {
            prev_byte = sm[(int)(32u * threadIdx.x) - 1];
            if (!(w.bnd & 1u) && prev_byte == '\\') {  // backslash parity in front of the warp's block: walk back over the run
              uint32_t nrun = 0, p = base;
              while (p > 0 && src.byte(p - 1) == '\\') {
                nrun++;
                p--;
                if ((bitmap[p >> 5] >> (p & 31u)) & 1u) break;
              }
              esc_in = (uint8_t)(nrun & 1u);
            }
          }
        } else {
          prev_byte = base <= n && base > 0 ? sm[(int)(32u * threadIdx.x) - 1] : (uint8_t)0;
          const int run = nlex2::clz32(~bs_prev);
          esc_in = (uint8_t)(run >= 32 ? 0 : (run & 1));  // (32 backslashes in a row: the statement is flagged below)
        }
        prev2_byte = base <= n && base > 1 ? sm[(int)(32u * threadIdx.x) - 2] : (uint8_t)0;
      }
      const uint32_t escm = nlex2::esc_mask32(w.bs, esc_in) & ~w.bnd;
      const nlex2::Events ev = nlex2::make_events(w, escm, prev_byte);
      uint32_t fnv = NUTDB_VEC8_ID;
      if (base < n)
        fnv = ev.all ? nlex2::ctx_window_fn(S.T, w, ev, NUTDB_VEC8_ID)
                     : vec8_then_row(NUTDB_VEC8_ID, S.T.a_row[EV_OTHER][0], S.T.a_row[EV_OTHER][1]);
      uint32_t fexcl;
      {
        const uint32_t incl = warp_scan_vec8(fnv, lane, fexcl);
        if (lane == 31) S.wfn[warp] = incl;
      }
      if (threadIdx.x == 0) {
        const uint8_t b0q = (w.sq & 1u) ? (uint8_t)'\'' : ((w.dq & 1u) ? (uint8_t)'"' : (uint8_t)0);
        S.twin_entry = (b0q && prev_byte == b0q && !(w.bnd & 1u)) ? 1u : 0u;
      }
      if (threadIdx.x == 0)  // the byte behind the tile starts a statement (or is the end of the batch)
        S.bndm[L3_WIN] = (tile_begin + L3_TILE < n ? (bm[L3_WIN] & 1u) : 0u) | (tile_begin + L3_TILE == n ? 1u : 0u);
      S.bndm[threadIdx.x] = w.bnd;
      L3_STAMP(2);
      bar_workers();
      if (threadIdx.x == 0) {
        // The tile's own function is published at once -- successors compose it without waiting for OUR look-back.
        // A statement start in the tile makes it a constant function: that already is the inclusive state.
        uint32_t agg = NUTDB_VEC8_ID;
#pragma unroll
        for (int i = 0; i < L3_WARPS; i++) agg = vec8_then(agg, S.wfn[i]);
        st_pub_u64(desc.fn + tile, ((unsigned long long)((epoch << 2) | (vec8_is_const(agg) ? 2u : 1u)) << 32) | agg);
      }
      bar_all();  // ---- A: wfn complete; the helper has the tile's entry state
      L3_STAMP(3);
      uint8_t s_warp;
      {
        uint32_t st = S.s_tile_in;
        for (int i = 0; i < warp; i++) st = vec8_apply(S.wfn[i], st);
        s_warp = (uint8_t)st;
      }
      const uint8_t s_in = (uint8_t)vec8_apply(fexcl, s_warp);
      // ---------------- stage 2: concrete context walk, token masks ----------------
      if (live) nlex3::ctx_window3(w, ev, base, nx, s_in, prev_byte, o);
      o.escm = escm;
      nlex3::Hist3 h;
      {
        h.L = __shfl_up_sync(full, w.L & o.ct, 1);
        h.D = __shfl_up_sync(full, w.D & o.ct, 1);
        h.DOT = __shfl_up_sync(full, w.DOT & o.ct, 1);
        h.bnd = __shfl_up_sync(full, w.bnd, 1);
        // lane 0: the 32 bytes in front of the warp's block; their raw classes are exact where it matters (a run of
        // word characters / dots that reaches the block) when the block is entered in code
        uint32_t rL = 0, rD = 0, rDOT = 0;
        if (blk >= 32u) {
          const uint8_t c = sm[(int)(blk - tile_begin) - 32 + lane];
          const uint8_t pr = S.T.prop[c];
          rL = __ballot_sync(full, (pr & PR_WORD) && !(pr & PR_DIGIT));
          rD = __ballot_sync(full, pr & PR_DIGIT);
          rDOT = __ballot_sync(full, c == '.');
        }
        if (lane == 0) {
          const bool code_entry = s_warp <= A_CX;
          h.L = code_entry ? rL : 0u;
          h.D = code_entry ? rD : 0u;
          h.DOT = code_entry ? rDOT : 0u;
          h.bnd = blk >= 32u ? (warp == 0 ? bitmap[(blk - 32u) >> 5] : bm[(blk - 32u - tile_begin) >> 5]) : 0u;
        }
      }
      uint32_t bad_carry = 0;
      if (live) {
        nlex3::win_tokens3(w, op, o, h, nx, prev_byte, prev2_byte, m);
        uint32_t bad = m.bad | nlex2::win_bad_mask(src, w, o, base, prev_byte);
        if (w.bs == 0xFFFFFFFFu) bad |= 1u;  // a backslash run longer than a window: parity not tracked
        // flag statements: find the start of the statement each flagged byte belongs to
        // (the statement open at the window start is only known after look-back (b): remember the request)
        uint32_t bb = bad;
        const uint32_t bnds = w.bnd & w.valid;
        while (bb) {
          const int i = __ffs((int)bb) - 1;
          bb &= bb - 1;
          const uint32_t below = bnds & (i >= 31 ? 0xFFFFFFFFu : ((2u << i) - 1u));
          if (below) out.punt_stmt_at(base + (uint32_t)(31 - __clz((int)below)));
          else bad_carry = 1;
        }
        bb = o.bad_prev;  // the statement that ENDS right before this (statement start) byte
        while (bb) {
          const int i = __ffs((int)bb) - 1;
          bb &= bb - 1;
          const uint32_t below = bnds & ((1u << i) - 1u);
          if (below) out.punt_stmt_at(base + (uint32_t)(31 - __clz((int)below)));
          else bad_carry = 1;
        }
        // the batch ends exactly on a window boundary inside a literal / comment: no window carries the virtual start
        if (base + 32u == n && (o.s_out == A_SQ || o.s_out == A_DQ || o.s_out == A_BT || o.s_out == A_BC0 || o.s_out == A_BC)) {
          if (bnds) out.punt_stmt_at(base + (uint32_t)(31 - __clz((int)bnds)));
          else bad_carry = 1;
        }
        mine.x = (uint32_t)__popc(m.has) + (uint32_t)__popc(m.eofm);
        mine.y = o.last_bnd1;
        mine.z = o.sc.open_pos;
        mine.w = (uint32_t)(o.sc.has_open != 0) | ((uint32_t)(o.sc.esc != 0) << 1) | ((uint32_t)(o.sc.chk != 0) << 3);
        // a literal opened in an earlier window closes here: its record needs look-back (b)'s open-quote offset
        deferred = nlex3::has_carried_close(o);
      }
      // ---------------- scan of (count, statement start, open literal) over the tile ----------------
      uint4 incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint4 o2 = c3_shfl_up(incl, d);
        if (lane >= d) incl = c3_then(o2, incl);
      }
      excl = c3_shfl_up(incl, 1);
      if (lane == 0) excl = make_uint4(0u, 0u, 0u, 0u);
      if (lane == 31) S.wsum[warp] = incl;
      if (bad_carry) mine.w |= 4u;
      L3_STAMP(4);
      bar_workers();  // ---- C: wsum complete
      L3_STAMP(5);
      if (threadIdx.x == 0) {
        __threadfence_block();
        S.agg_ready = tile + 1u;  // the helper takes it from here (publishes the sums, finishes look-back (b))
      }
      {
        uint4 pre = make_uint4(0u, 0u, 0u, 0u);
        for (int i = 0; i < warp; i++) pre = c3_then(pre, S.wsum[i]);
        excl = c3_then(pre, excl);  // everything of this tile before the thread's window
      }
      local = excl.x;
    }
    // ---------------- stage 3: token records (thread per window) -> tokens (thread per token) ----------------
    uint32_t tile_count;
    {
      uint32_t t = 0;
      for (int i = 0; i < L3_WARPS; i++) t += S.wsum[i].x;
      tile_count = t;  // (the helper reads wsum after agg_ready, the workers after barrier C)
    }
    for (uint32_t r0 = 0; r0 < tile_count; r0 += L3_RCAP) {
      auto rec = [&](uint32_t idx, uint32_t start_abs, uint32_t end_abs, uint32_t flags) {
        const uint32_t k = idx - r0;
        if (k < L3_RCAP) {
          S.rec0[k] = start_abs;
          S.rec1[k] = (end_abs - tile_begin) | (flags << NUTDB_R3_KIND_SHIFT);
        }
      };
      const bool mine_now = live && mine.x && local < r0 + L3_RCAP && local + mine.x > r0;
      if (r0) bar_all();  // the previous round's records have been consumed
      // windows whose records do not depend on look-back (b) write them while the helper finishes it
      nlex2::StrCarry none;
      if (mine_now && !deferred) nlex3::win_records3(o, m, base, none, local, rec);
      if (r0 == 0) {
        L3_STAMP(helper ? 4 : 6);
        bar_all();  // ---- D: look-back (b) done: c_tile_pre
        L3_STAMP(helper ? 5 : 7);
        if (!helper) {
          const uint4 cin = c3_then(S.c_tile_pre, excl);  // everything before this thread's window
          const uint32_t sst_open = cin.y ? cin.y - 1u : 0u;
          if (out.dbg && threadIdx.x == 0) {
            out.dbg[4 * tile + 0] = S.s_tile_in;
            out.dbg[4 * tile + 1] = cin.x;
            out.dbg[4 * tile + 2] = cin.y;
            out.dbg[4 * tile + 3] = cin.z | (cin.w << 30);
          }
          S.sst_in[threadIdx.x] = sst_open;
          if (live && (mine.w & 4u)) out.punt_stmt_at(sst_open);
          if (base <= n) {  // (the window that starts exactly at the batch end holds the last statement's token-range end)
            out.win_idx[base >> 5] = cin.x;
            out.win_has[base >> 5] = m.has;
            out.win_eof[base >> 5] = m.eofm;
          }
          // keep the open-literal carry for the deferred windows
          mine.z = cin.z;
          mine.w = (mine.w & ~11u) | (cin.w & 11u);
        }
      }
      if (mine_now && deferred) {
        nlex2::StrCarry sc_in;
        sc_in.has_open = (uint8_t)(mine.w & 1u);
        sc_in.esc = (uint8_t)((mine.w >> 1) & 1u);
        sc_in.chk = (uint8_t)((mine.w >> 3) & 1u);
        sc_in.open_pos = mine.z;
        nlex3::win_records3(o, m, base, sc_in, local, rec);
      }
      bar_all();  // ---- E: records complete
      if (r0 == 0) L3_STAMP(helper ? 6 : 8);
      const uint32_t tile_first = S.c_tile_pre.x;  // index of the tile's first token
      const uint32_t cnt = min(tile_count - r0, (uint32_t)L3_RCAP);
      for (uint32_t k = threadIdx.x; k < cnt; k += L3_THREADS) {
        const uint32_t start_abs = S.rec0[k], r1 = S.rec1[k];
        const uint32_t end_rel = r1 & ((1u << NUTDB_R3_KIND_SHIFT) - 1u), flags = r1 >> NUTDB_R3_KIND_SHIFT;
        const uint32_t last = end_rel - 1u;  // the token's last byte, tile relative
        const uint32_t wv = last >> 5, i = last & 31u;
        const uint32_t bb = S.bndm[wv] & (i >= 31u ? 0xFFFFFFFFu : ((2u << i) - 1u));
        const uint32_t sst = bb ? tile_begin + 32u * wv + (uint32_t)(31 - __clz((int)bb)) : S.sst_in[wv];
        nlex3::Tok3 tk;
        if (nlex3::token_finish3(S.T, src, start_abs, tile_begin + end_rel, flags, sst, tk))
          tk.kw = nlex3::token_keyword3(S.T, src, start_abs, tile_begin + end_rel - start_abs);
        if (tk.punt) {
          out.punt_stmt_at(sst);
          tk.type = NUTDB_TT_POISON;  // (the slot belongs to a flagged statement: a fixed filler)
          tk.start = tk.end = 0;
          tk.kw = 0;
        }
        const uint32_t gi = tile_first + r0 + k;
        if (gi < out.cap) {
          out.type[gi] = tk.type;
          out.start[gi] = tk.start;
          out.end[gi] = tk.end;
          out.kw[gi] = tk.kw;
        }
      }
    }
    if (tile_count == 0) {  // (no round ran: the barriers D and E, and the per-window outputs)
      bar_all();
      if (!helper) {
        const uint4 cin = c3_then(S.c_tile_pre, excl);
        if (out.dbg && threadIdx.x == 0) {
          out.dbg[4 * tile + 0] = S.s_tile_in;
          out.dbg[4 * tile + 1] = cin.x;
          out.dbg[4 * tile + 2] = cin.y;
          out.dbg[4 * tile + 3] = cin.z | (cin.w << 30);
        }
        if (live && (mine.w & 4u)) out.punt_stmt_at(cin.y ? cin.y - 1u : 0u);
        if (base <= n) {
          out.win_idx[base >> 5] = cin.x;
          out.win_has[base >> 5] = m.has;
          out.win_eof[base >> 5] = m.eofm;
        }
      }
    }
    if (b) phase1 ^= 1u;
    else phase0 ^= 1u;
    L3_STAMP(helper ? 7 : 9);
    bar_all();  // ---- F: everyone is done with buffer b and the per-tile tables before the next tile reuses them
  }
}
