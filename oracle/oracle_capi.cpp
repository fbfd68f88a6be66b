// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_lex.hpp header).
// C API over the restated reference parser, loaded with ctypes by tests/ and bench.py.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "oracle_fmt.hpp"

using namespace ora;

namespace {

struct OneResult {
  int status = 0;
  std::string debug;
  std::string error;
  NutdbError rec{};
  std::vector<NutdbNode> nodes;
  std::vector<PulledToken> pulled;
  size_t m_alg = 0;
};

void fill_error_record(const ParseErr& e, NutdbError& r) {
  std::memset(&r, 0, sizeof(r));
  if (e.is_lex) {
    r.cls = NUTDB_ST_LEX_ERROR;
    r.code = (uint16_t)e.lex.site;
    r.line = (uint32_t)e.lex.pos.line;
    r.col = (uint32_t)e.lex.pos.col;
    r.pos = (uint32_t)e.lex.byte_pos;
  } else {
    r.cls = NUTDB_ST_SYNTAX_ERROR;
    r.code = (uint16_t)e.syn.variant;
    if (e.syn.has_pos) {
      r.line = (uint32_t)e.syn.pos.line;
      r.col = (uint32_t)e.syn.pos.col;
      r.pos = (uint32_t)e.syn.byte_pos;
    }
    r.a = e.syn.a;
    r.b = e.syn.b;
    r.c = e.syn.c;
  }
}

// Parser::parse (mod.rs:27); `want` bit0: debug string, bit1: flat nodes, bit2: error display
void parse_one(sv sql, int want, OneResult& out) {
  Parser p(sql);
  try {
    Statement st = p.parse_stmt();
    out.status = NUTDB_ST_OK;
    if (want & 1) {
      DebugFmt f;
      f.statement(st);
      out.debug = std::move(f.o);
    }
    if (want & 2) {
      Flat fl;
      fl.base = sql.data();
      fl.statement(st);
      out.nodes = std::move(fl.n);
      out.m_alg = fl.m_alg;
    }
  } catch (const ParseErr& e) {
    out.status = e.is_lex ? NUTDB_ST_LEX_ERROR : NUTDB_ST_SYNTAX_ERROR;
    fill_error_record(e, out.rec);
    if (want & 4) out.error = error_display(e);
  } catch (const ReferencePanic&) {
    out.status = NUTDB_ST_REFERENCE_PANIC;
    std::memset(&out.rec, 0, sizeof(out.rec));
    out.rec.cls = NUTDB_ST_REFERENCE_PANIC;
    if (want & 4) out.error = "the reference panics on this input (unreachable!() at literal.rs:63)";
  }
  out.pulled = std::move(p.pulled);
}

}  // namespace

extern "C" {

// Full token stream of the reference tokenizer (incl. Whitespace, Comment and the final EOF)
// up to the first error.  Returns the number of tokens; *err_site != 0 if an error ended it.
int64_t ora_tokenize(const char* sql, size_t len, uint8_t* types, uint32_t* starts, uint32_t* ends, size_t cap,
                     int* err_type, int* err_site, uint32_t* err_pos, uint32_t* err_line, uint32_t* err_col,
                     char* ctx, size_t ctxcap) {
  Tokenizer t(sv(sql, len));
  int64_t n = 0;
  *err_site = 0;
  for (;;) {
    TokResult r = t.next_token();
    if (!r.ok) {
      *err_type = (int)r.err.t;
      *err_site = r.err.site;
      *err_pos = (uint32_t)r.err.byte_pos;
      *err_line = (uint32_t)r.err.pos.line;
      *err_col = (uint32_t)r.err.pos.col;
      if (ctx && ctxcap) {
        size_t m = std::min(ctxcap - 1, r.err.ctx.size());
        std::memcpy(ctx, r.err.ctx.data(), m);
        ctx[m] = 0;
      }
      return n;
    }
    if ((size_t)n < cap) {
      types[n] = r.tok.t;
      starts[n] = (uint32_t)r.tok.span.start;
      ends[n] = (uint32_t)r.tok.span.end;
    }
    n++;
    if (r.tok.t == NUTDB_TT_EOF) return n;
  }
}

void ora_get_pos(const char* sql, size_t len, size_t cursor, uint32_t* line, uint32_t* col) {
  Utf8Iter it(sv(sql, len));
  Position p = it.get_pos(cursor);
  *line = (uint32_t)p.line;
  *col = (uint32_t)p.col;
}

// literal.rs unescape_{single,double}_quoted_string; returns 0 ok, 1 InvalidEscapedUnicode
int ora_unescape(const char* raw, size_t len, int quote, char* out, size_t cap, size_t* outlen) {
  std::string res, bad;
  Span hs;
  bool ok;
  try {
    ok = unescape_string(sv(raw, len), (char32_t)quote, res, bad, hs);
  } catch (const ReferencePanic&) {
    *outlen = 0;
    return 2;
  }
  const std::string& s = ok ? res : bad;
  *outlen = s.size();
  std::memcpy(out, s.data(), std::min(cap, s.size()));
  return ok ? 0 : 1;
}

int ora_keyword_id(const char* s, size_t len) { return keyword_id(sv(s, len)); }
const char* ora_keyword_text(int id) { return (id >= 1 && id <= NUTDB_KW_COUNT) ? KEYWORDS[id - 1] : ""; }

void* ora_parse(const char* sql, size_t len) {
  auto* r = new OneResult();
  parse_one(sv(sql, len), 7, *r);
  return r;
}
void ora_free(void* h) { delete (OneResult*)h; }
int ora_status(void* h) { return ((OneResult*)h)->status; }
const char* ora_debug(void* h) { return ((OneResult*)h)->debug.c_str(); }
const char* ora_error_display(void* h) { return ((OneResult*)h)->error.c_str(); }
void ora_error_record(void* h, NutdbError* out) { *out = ((OneResult*)h)->rec; }
size_t ora_n_nodes(void* h) { return ((OneResult*)h)->nodes.size(); }
const NutdbNode* ora_nodes(void* h) { return ((OneResult*)h)->nodes.data(); }
size_t ora_m_alg(void* h) { return ((OneResult*)h)->m_alg; }
size_t ora_n_pulled(void* h) { return ((OneResult*)h)->pulled.size(); }
void ora_pulled(void* h, uint8_t* types, uint32_t* starts, uint32_t* ends) {
  auto& p = ((OneResult*)h)->pulled;
  for (size_t i = 0; i < p.size(); i++) {
    types[i] = p[i].t;
    starts[i] = (uint32_t)p[i].span.start;
    ends[i] = (uint32_t)p[i].span.end;
  }
}

// ------------------------------------------------------------------------------------------
// batch: flat outputs for bit-exact comparison with the GPU, laid out like NutdbBatch
// ------------------------------------------------------------------------------------------
struct OraBatch {
  std::vector<NutdbStmt> stmt;
  std::vector<NutdbNode> node;
  std::vector<NutdbError> err;
  std::vector<uint8_t> tok_type, tok_kw;  // tok_kw: keyword id of a KeywordOrIdentifier token (keyword.rs), else 0
  std::vector<uint32_t> tok_start, tok_end;
  uint64_t t_alg = 0, m_alg = 0;
};

void* ora_parse_batch(const char* text, const uint64_t* offs, uint64_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::vector<OneResult> res(n);
  std::atomic<uint64_t> next{0};
  auto work = [&]() {
    for (;;) {
      uint64_t lo = next.fetch_add(256);
      if (lo >= n) return;
      uint64_t hi = std::min(n, lo + 256);
      for (uint64_t i = lo; i < hi; i++) parse_one(sv(text + offs[i], offs[i + 1] - offs[i]), 2, res[i]);
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();

  auto* b = new OraBatch();
  b->stmt.resize(n);
  uint64_t nn = 0, nt = 0;
  for (uint64_t i = 0; i < n; i++) {
    nn += res[i].nodes.size();
    nt += res[i].pulled.size();
  }
  b->node.reserve(nn);
  b->tok_type.reserve(nt);
  b->tok_kw.reserve(nt);
  b->tok_start.reserve(nt);
  b->tok_end.reserve(nt);
  for (uint64_t i = 0; i < n; i++) {
    NutdbStmt& s = b->stmt[i];
    s.status = (uint32_t)res[i].status;
    s.tok_begin = (uint32_t)b->tok_type.size();
    s.tok_count = (uint32_t)res[i].pulled.size();
    s.tok_used = s.tok_count;
    s.node_begin = (uint32_t)b->node.size();
    s.node_count = (uint32_t)res[i].nodes.size();
    b->node.insert(b->node.end(), res[i].nodes.begin(), res[i].nodes.end());
    for (auto& p : res[i].pulled) {
      b->tok_type.push_back(p.t);
      b->tok_start.push_back((uint32_t)p.span.start);
      b->tok_end.push_back((uint32_t)p.span.end);
      b->tok_kw.push_back(p.t == NUTDB_TT_KeywordOrIdentifier
                              ? (uint8_t)keyword_id(sv(text + offs[i] + p.span.start, p.span.end - p.span.start))
                              : (uint8_t)0);
    }
    if (res[i].status != NUTDB_ST_OK) {
      NutdbError e = res[i].rec;
      e.stmt = (uint32_t)i;
      b->err.push_back(e);
    }
    b->t_alg += res[i].pulled.size();
    b->m_alg += res[i].m_alg;
  }
  return b;
}
void ora_batch_free(void* h) { delete (OraBatch*)h; }
void ora_batch_counts(void* h, uint64_t* n_node, uint64_t* n_err, uint64_t* n_tok, uint64_t* t_alg, uint64_t* m_alg) {
  auto* b = (OraBatch*)h;
  *n_node = b->node.size();
  *n_err = b->err.size();
  *n_tok = b->tok_type.size();
  *t_alg = b->t_alg;
  *m_alg = b->m_alg;
}
const NutdbStmt* ora_batch_stmt(void* h) { return ((OraBatch*)h)->stmt.data(); }
const NutdbNode* ora_batch_node(void* h) { return ((OraBatch*)h)->node.data(); }
const NutdbError* ora_batch_err(void* h) { return ((OraBatch*)h)->err.data(); }
const uint8_t* ora_batch_tok_type(void* h) { return ((OraBatch*)h)->tok_type.data(); }
const uint8_t* ora_batch_tok_kw(void* h) { return ((OraBatch*)h)->tok_kw.data(); }
const uint32_t* ora_batch_tok_start(void* h) { return ((OraBatch*)h)->tok_start.data(); }
const uint32_t* ora_batch_tok_end(void* h) { return ((OraBatch*)h)->tok_end.data(); }

// ------------------------------------------------------------------------------------------
// CPU baseline: parse + drop every statement of the batch `reps` times on `nthreads` threads
// (one contiguous byte-balanced range per thread, like one rayon/std::thread per core around
// Parser::parse).  Returns best-of-reps wall seconds; outputs statements ok/failed and pulled tokens.
// ------------------------------------------------------------------------------------------
double ora_bench(const char* text, const uint64_t* offs, uint64_t n, int nthreads, int reps, uint64_t* n_ok,
                 uint64_t* n_tokens) {
  if (nthreads < 1) nthreads = 1;
  // contiguous statement ranges balanced by bytes
  std::vector<uint64_t> cut(nthreads + 1, n);
  cut[0] = 0;
  uint64_t total = offs[n] - offs[0];
  {
    uint64_t s = 0;
    for (int t = 1; t < nthreads; t++) {
      uint64_t target = offs[0] + total * (uint64_t)t / (uint64_t)nthreads;
      while (s < n && offs[s] < target) s++;
      cut[t] = s;
    }
  }
  double best = 1e30;
  for (int rep = 0; rep < reps; rep++) {
    std::vector<uint64_t> ok(nthreads, 0), toks(nthreads, 0);
    auto work = [&](int t) {
      uint64_t o = 0, k = 0;
      for (uint64_t i = cut[t]; i < cut[t + 1]; i++) {
        Parser p(sv(text + offs[i], offs[i + 1] - offs[i]));
        try {
          Statement st = p.parse_stmt();
          o++;
        } catch (const ParseErr&) {
        } catch (const ReferencePanic&) {
        }
        k += p.pulled.size();
      }
      ok[t] = o;
      toks[t] = k;
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    if (s < best) best = s;
    uint64_t so = 0, sk = 0;
    for (int t = 0; t < nthreads; t++) {
      so += ok[t];
      sk += toks[t];
    }
    *n_ok = so;
    *n_tokens = sk;
  }
  return best;
}

int ora_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
