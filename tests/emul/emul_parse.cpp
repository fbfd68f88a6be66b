// TEST HARNESS (not shipped): runs the device lexer + parser logic (lex_core.cuh, parse_core.cuh)
// on the host, statement by statement, producing the same arrays the CUDA path produces.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../nutdb_b200/csrc/lex_tables.hpp"
#include "../../nutdb_b200/csrc/parse_fast.cuh"
#include "../../nutdb_b200/csrc/parse_fast_tables.hpp"

using namespace nlex;

extern "C" int64_t emul_lex(const uint8_t* text, uint32_t n, const uint64_t* offs, uint64_t nstmt, int emit_all,
                            uint32_t chunk, uint8_t* tok_type, uint32_t* tok_start, uint32_t* tok_end, uint8_t* tok_kw,
                            uint32_t cap, uint32_t* seg_tok_begin, uint32_t* seg_tok_end, uint32_t* n_seg_out);

extern "C" int64_t emul_lex2(const uint8_t* text, uint32_t n, const uint64_t* offs, uint64_t nstmt, uint32_t seg_len,
                             uint8_t* tok_type, uint32_t* tok_start, uint32_t* tok_end, uint8_t* tok_kw, uint32_t cap,
                             uint32_t* stmt_tok_begin, uint32_t* stmt_tok_end);

extern "C" int64_t emul_lex3(const uint8_t* text, uint32_t n, const uint64_t* offs, uint64_t nstmt, uint32_t seg_len,
                             uint8_t* tok_type, uint32_t* tok_start, uint32_t* tok_end, uint8_t* tok_kw, uint32_t cap,
                             uint32_t* stmt_tok_begin, uint32_t* stmt_tok_end);

namespace {
struct HTok {
  const uint8_t* ty;
  const uint32_t* st;
  const uint32_t* en;
  const uint8_t* kw;
  uint32_t n;
  uint8_t type(uint32_t i) const { return i < n ? ty[i] : (uint8_t)NUTDB_TT_EOF; }
  uint8_t kwid(uint32_t i) const { return i < n ? kw[i] : 0; }
  uint32_t start(uint32_t i) const { return i < n ? st[i] : 0; }
  uint32_t end(uint32_t i) const { return i < n ? en[i] : 0; }
};
struct HTokAdapter {
  HTok h;
  uint8_t type(uint32_t i) const { return h.type(i); }
  uint8_t kw(uint32_t i) const { return h.kwid(i); }
  uint8_t type_at(uint32_t i) const { return h.type(i); }
  uint8_t kw_at(uint32_t i) const { return h.kwid(i); }
  uint32_t pair_at(uint32_t i) const { return (uint32_t)h.type(i) | ((uint32_t)h.kwid(i) << 8); }
  uint32_t start(uint32_t i) const { return h.start(i); }
  uint32_t end(uint32_t i) const { return h.end(i); }
};
struct HNodes {
  std::vector<npar::CNode>& v;
  uint32_t cap;
  npar::CNode get(uint32_t i) const { return v[i]; }
  void set(uint32_t i, const npar::CNode& x) {
    if (i >= v.size()) v.resize(i + 1);
    v[i] = x;
  }
  void set_raw(uint32_t i, uint32_t header, uint32_t x) {
    npar::CNode c;
    c.kind = (uint8_t)(header & 0xFF);
    c.sub = (uint8_t)((header >> 8) & 0xFF);
    c.aux = (uint16_t)(header >> 16);
    c.x = x;
    set(i, c);
  }
  uint32_t capacity() const { return cap; }
};
struct HCompact {
  const std::vector<npar::CNode>& v;
  npar::CNode operator()(uint32_t i) const { return v[i]; }
};
struct HOut {
  NutdbNode* o;
  void body(uint32_t j, uint8_t kind, uint8_t sub, uint16_t aux, uint32_t a, uint32_t b) {
    o[j].kind = kind;
    o[j].sub = sub;
    o[j].aux = aux;
    o[j].a = a;
    o[j].b = b;
  }
  void parent(uint32_t j, uint32_t p) { o[j].parent = p; }
};
struct HText {
  const uint8_t* p;
  uint32_t n;
  uint8_t byte(uint32_t i) const { return i < n ? p[i] : 0; }
  uint8_t raw(uint32_t i) const { return p[i]; }
  // the device's DText::skip_plain: same four-byte granularity and the same zero-byte test on the word (the device
  // builds the word from two aligned loads); checked here against the plain byte comparison on every call
  uint32_t skip_plain(uint32_t p0, uint32_t e, uint32_t quote) const {
    const uint32_t q4 = quote * 0x01010101u;
    while (p0 + 4u <= e) {
      bool hit = false;
      for (uint32_t k = 0; k < 4; k++) hit |= p[p0 + k] == '\\' || p[p0 + k] == quote;
      uint32_t v;
      std::memcpy(&v, p + p0, 4);  // (little-endian host, like the device)
      const uint32_t x = v ^ 0x5C5C5C5Cu, y = v ^ q4;
      const bool swar = ((((x - 0x01010101u) & ~x) | ((y - 0x01010101u) & ~y)) & 0x80808080u) != 0u;
      if (swar != hit) std::abort();
      if (hit) break;
      p0 += 4u;
    }
    return p0;
  }
};
}  // namespace

static int g_use_fast = 2;        // 0 = automaton only, 1 = narrow table-driven pass first, 2 = narrow, then wide
static int g_lexer = 3;          // 1 = thread-per-chunk walker, 2 = three-pass mask lexer, 3 = single-pass lexer
static uint32_t g_seg_len = 1024;
static uint64_t g_fast_hits = 0, g_wide_hits = 0;

extern "C" {

void emul_set_fast(int on) { g_use_fast = on; }
void emul_set_lexer(int version, uint32_t seg_len) {
  g_lexer = version;
  g_seg_len = seg_len;
}
uint64_t emul_fast_hits(void) { return g_fast_hits; }
uint64_t emul_wide_hits(void) { return g_wide_hits; }
void emul_reset_fast_hits(void) { g_fast_hits = g_wide_hits = 0; }

// Parses every statement.  Outputs: stmt[nstmt] (NutdbStmt), nodes (compact, caller cap), errors.
// Returns 0, or -1 on capacity overflow.
int emul_parse_batch(const uint8_t* text, const uint64_t* offs, uint64_t nstmt, uint32_t chunk, uint32_t stack_cap,
                     NutdbStmt* stmt, NutdbNode* nodes, uint64_t node_cap, uint64_t* n_nodes, NutdbError* errs,
                     uint64_t err_cap, uint64_t* n_errs, uint8_t* tok_type, uint32_t* tok_start, uint32_t* tok_end,
                     uint8_t* tok_kw, uint32_t tok_cap, uint64_t* n_tok) {
  uint32_t n = (uint32_t)(offs[nstmt] - offs[0]);
  const uint8_t* base = text + offs[0];
  std::vector<uint32_t> sb(nstmt + 1), se(nstmt + 1);
  uint32_t nseg = 0;
  int64_t nt;
  if (g_lexer == 3) {
    nt = emul_lex3(base, n, offs, nstmt, g_seg_len, tok_type, tok_start, tok_end, tok_kw, tok_cap, sb.data(), se.data());
  } else if (g_lexer == 2) {
    nt = emul_lex2(base, n, offs, nstmt, g_seg_len, tok_type, tok_start, tok_end, tok_kw, tok_cap, sb.data(), se.data());
  } else {
    nt = emul_lex(base, n, offs, nstmt, 0, chunk, tok_type, tok_start, tok_end, tok_kw, tok_cap, sb.data(), se.data(), &nseg);
  }
  if (nt < 0) return (int)nt;
  *n_tok = (uint64_t)nt;
  uint64_t nn = 0, ne = 0;
  uint32_t seg = 0;
  std::vector<uint32_t> stack(stack_cap);
  std::vector<npar::CNode> tmp;
  for (uint64_t s = 0; s < nstmt; s++) {
    uint32_t len = (uint32_t)(offs[s + 1] - offs[s]);
    NutdbStmt& S = stmt[s];
    std::memset(&S, 0, sizeof(S));
    npar::ParseResult res;
    HText tx{base + (offs[s] - offs[0]), len};
    tmp.clear();
    if (len == 0) {
      // Parser::parse("") : the first token is EOF => EmptyQuery (mod.rs:141-144)
      S.status = NUTDB_ST_SYNTAX_ERROR;
      S.tok_begin = (uint32_t)nt;
      S.tok_count = 0;
      S.tok_used = 1;
      res.status = NUTDB_ST_SYNTAX_ERROR;
      res.err_code = NUTDB_SE_EmptyQuery;
      res.err_has_pos = false;
      res.err_pos = res.err_a = res.err_b = res.err_c = 0;
    } else {
      // lexer 1 indexes token ranges by non-empty statement ordinal, lexer 2 by statement
      uint32_t b = g_lexer >= 2 ? sb[s] : sb[seg], e = g_lexer >= 2 ? se[s] : se[seg];
      seg++;
      HTokAdapter tk{HTok{tok_type + b, tok_start + b, tok_end + b, tok_kw + b, e - b}};
      HNodes nd{tmp, 2 * (e - b) + 8};
      // same order as the device: the straight-line parser first, the exact automaton if it declines
      bool fast = false;
      if (g_use_fast) {
        static npar::FastTables FT;
        static bool ft_init = false;
        if (!ft_init) {
          npar::fast_tables_build(FT);
          ft_init = true;
        }
        npar::FastStackEntry fstack[FAST_STACK_DEPTH];
        npar::FastParser<HTokAdapter, HNodes, HText> f(&FT, tk, nd, tx, fstack, 1);
        fast = f.try_parse(res);
        if (!fast && g_use_fast >= 2) {
          // the wide pass: nodes and operator stack share one range of tok_count + NODE_SLACK slots (nodes from the
          // bottom, the stack from the top), exactly the device's layout
          const uint32_t wcap = (e - b) + 8;
          std::vector<npar::CNode> range(wcap);
          struct WNodes {
            npar::CNode* p;
            uint32_t cap;
            npar::CNode get(uint32_t i) const { return p[i]; }
            void set(uint32_t i, const npar::CNode& x) { p[i] = x; }
            void set_raw(uint32_t i, uint32_t header, uint32_t x) {
              npar::CNode c;
              c.kind = (uint8_t)(header & 0xFF);
              c.sub = (uint8_t)((header >> 8) & 0xFF);
              c.aux = (uint16_t)(header >> 16);
              c.x = x;
              p[i] = c;
            }
            uint32_t capacity() const { return cap; }
          } wn{range.data(), wcap};
          static_assert(sizeof(npar::CNode) == sizeof(npar::FastStackEntry), "the wide stack shares the node range");
          npar::FastParser<HTokAdapter, WNodes, HText, true> w(&FT, tk, wn, tx,
                                                                reinterpret_cast<npar::FastStackEntry*>(range.data()) + (wcap - 1), -1);
          fast = w.try_parse(res);
          if (fast) {
            g_wide_hits++;
            tmp.assign(range.begin(), range.begin() + res.node_count);
          }
        }
      }
      if (fast) {
        g_fast_hits++;
      } else {
        npar::Machine<HTokAdapter, HNodes, HText> m(npar::PARSE_TABLES, tk, nd, tx, stack.data(), stack_cap, res);
        m.run(NUTDB_PROGRAM_ENTRY);
      }
      S.status = res.status;
      S.tok_begin = b;
      S.tok_count = e - b;
      S.tok_used = res.tok_used;
    }
    if (res.status == NUTDB_ST_OK) {
      if (nn + res.node_count > node_cap) return -1;
      S.node_begin = (uint32_t)nn;
      S.node_count = res.node_count;
      {  // same per-node expansion the device runs with one thread per node
        uint32_t b = S.tok_begin;
        HTokAdapter tk{HTok{tok_type + b, tok_start + b, tok_end + b, tok_kw + b, S.tok_count}};
        HCompact cn{tmp};
        HOut out{nodes + nn};
        for (uint32_t j = 0; j < res.node_count; j++) npar::expand_node(cn, j, res.node_count, tk, out);
      }
      nn += res.node_count;
    } else {
      S.node_begin = (uint32_t)nn;
      S.node_count = 0;
      if (ne >= err_cap) return -1;
      NutdbError& E = errs[ne++];
      std::memset(&E, 0, sizeof(E));
      E.stmt = (uint32_t)s;
      E.cls = (uint16_t)res.status;
      E.code = res.err_code;
      E.a = res.err_a;
      E.b = res.err_b;
      E.c = res.err_c;
      if (res.err_has_pos) {
        E.pos = res.err_pos;
        npar::get_pos(tx, res.err_pos, E.line, E.col);
      }
    }
  }
  *n_nodes = nn;
  *n_errs = ne;
  return 0;
}

}  // extern "C"
