// TEST HARNESS (not shipped): drives the device lexer logic of nutdb_b200/csrc/lex_core.cuh on the
// host, with the same decomposition the CUDA kernel uses (fixed-size chunks; per-chunk transition
// functions of automata A and B; exclusive prefix composition; per-chunk count; per-chunk emit).
// Lets the CPU test-suite differential-test the lexer logic against the oracle for any chunk size.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../nutdb_b200/csrc/lex_tables.hpp"

using namespace nlex;

namespace {
struct HostSrc {
  const uint8_t* text;
  const uint32_t* bitmap;
  uint32_t n;
  uint8_t byte(uint32_t p) const { return p < n ? text[p] : 0; }
  bool boundary(uint32_t p) const { return p < n && ((bitmap[p >> 5] >> (p & 31)) & 1u); }
};
struct ArraySink {
  uint8_t* type;
  uint32_t* start;
  uint32_t* end;
  uint8_t* kw;
  uint32_t cap;
  uint32_t* seg_begin_arr;
  uint32_t* seg_end_arr;
  bool overflow = false;
  void token(uint32_t i, uint8_t t, uint32_t s, uint32_t e, uint8_t k) {
    if (i >= cap) { overflow = true; return; }
    type[i] = t;
    start[i] = s;
    end[i] = e;
    kw[i] = k;
  }
  void seg_begin(uint32_t seg, uint32_t first, uint32_t) { seg_begin_arr[seg] = first; }
  void seg_end(uint32_t seg, uint32_t endi, uint32_t) { seg_end_arr[seg] = endi; }
};
LexTables g_tables;
bool g_init = false;
}  // namespace

extern "C" {

const LexTables* emul_tables() {
  if (!g_init) {
    build_lex_tables(g_tables);
    g_init = true;
  }
  return &g_tables;
}

// Returns number of tokens, or -1 on capacity overflow.  seg arrays need nstmt entries.
int64_t emul_lex(const uint8_t* text, uint32_t n, const uint64_t* offs, uint64_t nstmt, int emit_all, uint32_t chunk,
                 uint8_t* tok_type, uint32_t* tok_start, uint32_t* tok_end, uint8_t* tok_kw, uint32_t cap,
                 uint32_t* seg_tok_begin, uint32_t* seg_tok_end, uint32_t* n_seg_out) {
  const LexTables& T = *emul_tables();
  std::vector<uint32_t> bitmap((n + 31) / 32 + 1, 0);
  uint32_t nseg = 0;
  for (uint64_t s = 0; s < nstmt; s++) {
    if (offs[s + 1] > offs[s]) {
      uint32_t p = (uint32_t)(offs[s] - offs[0]);
      bitmap[p >> 5] |= 1u << (p & 31);
      nseg++;
    }
  }
  *n_seg_out = nseg;
  if (n == 0) return 0;
  HostSrc src{text, bitmap.data(), n};
  uint32_t nchunks = (n + chunk - 1) / chunk;
  // phase A
  std::vector<uint32_t> vecA(nchunks), vecB(nchunks);
  std::vector<uint8_t> entA(nchunks), entB(nchunks);
  for (uint32_t c = 0; c < nchunks; c++) vecA[c] = chunk_sim_A(T, src, c * chunk, std::min(n, (c + 1) * chunk));
  uint32_t pref = NUTDB_VEC8_ID;
  for (uint32_t c = 0; c < nchunks; c++) {
    entA[c] = (uint8_t)vec8_apply(pref, A_C);
    pref = vec8_then(pref, vecA[c]);
  }
  // phase B
  for (uint32_t c = 0; c < nchunks; c++)
    vecB[c] = chunk_sim_B(T, src, c * chunk, std::min(n, (c + 1) * chunk), entA[c]);
  pref = NUTDB_VEC8_ID;
  for (uint32_t c = 0; c < nchunks; c++) {
    entB[c] = (uint8_t)vec8_apply(pref, B_N);
    pref = vec8_then(pref, vecB[c]);
  }
  // phase C
  std::vector<CSum> sums(nchunks), pre(nchunks);
  for (uint32_t c = 0; c < nchunks; c++)
    sums[c] = chunk_count(T, src, c * chunk, (c + 1) * chunk, n, entA[c], entB[c], emit_all != 0);
  CSum run = csum_identity();
  for (uint32_t c = 0; c < nchunks; c++) {
    pre[c] = run;
    run = csum_then(run, sums[c]);
  }
  // phase D
  ArraySink sink{tok_type, tok_start, tok_end, tok_kw, cap, seg_tok_begin, seg_tok_end};
  for (uint32_t c = 0; c < nchunks; c++) {
    if (emit_all) chunk_walk<true>(T, src, sink, c * chunk, (c + 1) * chunk, n, entA[c], entB[c], pre[c], false);
    else chunk_walk<false>(T, src, sink, c * chunk, (c + 1) * chunk, n, entA[c], entB[c], pre[c], false);
  }
  if (sink.overflow) return -1;
  return (int64_t)run.count;
}

uint8_t emul_keyword(const uint8_t* s, uint32_t len) {
  const LexTables& T = *emul_tables();
  return keyword_lookup(T, len, [s](uint32_t i) { return s[i]; });
}

}  // extern "C"
