// workload.cpp -- deterministic synthetic SQL batches for the benchmark configs of BASELINE.json
// (SURVEY.md section 8d): config 2 "short SELECT/INSERT/CREATE", config 3 "lexer stress"
// (string/comment/quoted-identifier heavy, 5 % malformed), config 4 "deep nesting".
//
// Bench/test tooling (libnutdb_workload.so): produces `text` + `stmt_off`, the inputs of
// nutdb_gpu_parse_batch.  Output depends only on (config, seed, target_bytes): statements are
// generated in blocks of WL_BLOCK statements, block k seeded with (seed, k), by any number of
// threads, and concatenated in block order.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr uint32_t WL_BLOCK = 2048;

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
  uint32_t range(uint32_t lo, uint32_t hi) { return lo + below(hi - lo + 1); }  // inclusive
  bool chance(uint32_t percent) { return below(100) < percent; }
};

const char* const KEYWORDS[] = {
    "by", "as", "on", "from", "intersect", "union", "all", "except", "distinct", "with", "select", "join", "where",
    "group", "having", "order", "limit", "offset", "using", "ties", "asc", "desc", "explain", "insert", "into",
    "values", "create", "primary", "key", "comment", "update", "default", "check", "describe", "drop", "alter", "add",
    "rename", "first", "after", "truncate", "optimize", "set", "database", "table", "view", "column", "index",
    "constraint", "partition", "null", "true", "false", "and", "or", "xor", "not", "in", "exists", "if", "case",
    "when", "then", "else", "end", "is", "between", "like", "ilike", "interval", "second", "minute", "hour", "day",
    "month", "year", "int8", "int16", "int32", "int64", "int128", "uint8", "uint16", "uint32", "uint64", "uint128",
    "serial32", "serial64", "serial128", "userial32", "userial64", "userial128", "decimal32", "decimal64", "float32",
    "float64", "boolean", "chars", "string", "uuid", "date", "datetime", "array", "enum", "tuple", "map",
    "dictionary", "nullable", "inner", "outer", "left", "right", "full", "semi", "anti"};

bool is_keyword(const std::string& w) {
  for (const char* k : KEYWORDS)
    if (w == k) return true;
  return false;
}

struct Gen {
  Rng r;
  std::string& o;
  Gen(uint64_t seed, std::string& out) : r(seed), o(out) {}

  // identifier [a-z_][a-z0-9_]{0,11}, never one of the 115 (contextual) keywords
  void ident() {
    static const char first[] = "abcdefghijklmnopqrstuvwxyz_";
    static const char rest[] = "abcdefghijklmnopqrstuvwxyz0123456789_";
    std::string w;
    w.push_back(first[r.below(27)]);
    uint32_t n = r.below(100) < 70 ? r.below(6) : r.below(12);
    for (uint32_t i = 0; i < n; i++) w.push_back(rest[r.below(37)]);
    if (is_keyword(w)) w.push_back('_');
    o += w;
  }
  void integer() { o += std::to_string(r.below(1000000)); }
  // 'string' of 0-16 printable ASCII chars; quote and backslash only through escapes
  void sq_string(bool allow_escape) {
    o.push_back('\'');
    uint32_t n = r.below(17);
    for (uint32_t i = 0; i < n; i++) {
      char c = (char)r.range(32, 126);
      if (c == '\'') {
        if (allow_escape) o += "''";
        else o.push_back(' ');
      } else if (c == '\\') {
        if (allow_escape) o += "\\\\";
        else o.push_back('/');
      } else {
        o.push_back(c);
      }
    }
    o.push_back('\'');
  }
  void literal() {
    uint32_t k = r.below(10);
    if (k < 6) integer();
    else if (k < 9) sq_string(false);
    else {
      integer();
      o.push_back('.');
      o += std::to_string(r.below(100));
    }
  }
  void cmp_op() {
    static const char* const ops[] = {"=", "!=", "<", "<=", ">", ">=", "<>"};
    o += ops[r.below(7)];
  }

  // ---------------------------------------------------------------- config 2
  void short_select() {
    o += "SELECT ";
    if (r.chance(8)) {
      o += "*";
    } else {
      uint32_t n = r.range(1, 5);
      for (uint32_t i = 0; i < n; i++) {
        if (i) o += ", ";
        uint32_t k = r.below(10);
        if (k == 0) {
          ident();
          o += "(";
          ident();
          o += ")";
        } else if (k == 1) {
          ident();
          o += " + ";
          integer();
        } else {
          ident();
        }
        if (r.chance(10)) {
          o += " AS ";
          ident();
        }
      }
    }
    o += " FROM ";
    ident();
    if (r.chance(65)) {
      o += " WHERE ";
      uint32_t n = r.range(1, 3);
      for (uint32_t i = 0; i < n; i++) {
        if (i) o += r.chance(75) ? " AND " : " OR ";
        ident();
        o.push_back(' ');
        cmp_op();
        o.push_back(' ');
        literal();
      }
    }
    if (r.chance(30)) {
      o += " ORDER BY ";
      ident();
      if (r.chance(50)) o += " DESC";
    }
    if (r.chance(30)) {
      o += " LIMIT ";
      o += std::to_string(r.range(1, 1000));
    }
  }
  void short_insert() {
    o += "INSERT INTO ";
    ident();
    uint32_t w = r.range(1, 5);
    o += " (";
    for (uint32_t i = 0; i < w; i++) {
      if (i) o += ", ";
      ident();
    }
    o += ") VALUES ";
    uint32_t rows = r.chance(70) ? 1 : r.range(2, 3);
    for (uint32_t j = 0; j < rows; j++) {
      if (j) o += ", ";
      o.push_back('(');
      for (uint32_t i = 0; i < w; i++) {
        if (i) o += ", ";
        literal();
      }
      o.push_back(')');
    }
  }
  void datatype(uint32_t depth) {
    static const char* const scalars[] = {"Int8", "Int16", "Int32", "Int64", "UInt8", "UInt16", "UInt32", "UInt64",
                                          "Float32", "Float64", "Boolean", "String", "UUID", "Date", "DateTime",
                                          "Int128", "Serial64"};
    uint32_t k = r.below(20);
    if (k == 0 && depth < 2) {
      o += "Array(";
      datatype(depth + 1);
      o += ")";
    } else if (k == 1 && depth < 2) {
      o += "Nullable(";
      datatype(depth + 1);
      o += ")";
    } else if (k == 2) {
      o += "Chars(";
      o += std::to_string(r.range(1, 64));
      o += ")";
    } else if (k == 3) {
      o += "Decimal64(";
      o += std::to_string(r.range(1, 18));
      o += ")";
    } else {
      o += scalars[r.below(17)];
    }
  }
  void short_create() {
    o += "CREATE TABLE ";
    if (r.chance(20)) o += "IF NOT EXISTS ";
    ident();
    o += " (";
    uint32_t n = r.range(1, 6);
    for (uint32_t i = 0; i < n; i++) {
      if (i) o += ", ";
      ident();
      o.push_back(' ');
      datatype(0);
      if (r.chance(10)) {
        o += " DEFAULT ";
        literal();
      }
    }
    o += ")";
    if (r.chance(30)) {
      o += " PRIMARY KEY ";
      ident();
    }
    if (r.chance(20)) {
      o += " ORDER BY ";
      ident();
    }
  }
  void config2() {
    uint32_t k = r.below(100);
    if (k < 60) short_select();
    else if (k < 85) short_insert();
    else short_create();
    o += ";\n";
  }

  // ---------------------------------------------------------------- config 3
  void utf8_char() {
    static const char* const pool[] = {"\xC3\xA9", "\xC3\xBC", "\xCE\xB2", "\xE4\xBD\xA0", "\xE5\xA5\xBD",
                                       "\xE2\x9D\xA4", "\xF0\x9F\x98\x80", "\xF0\x9F\x9A\x80"};
    o += pool[r.below(8)];
  }
  // body of a quoted literal: other quote kind, ';', "--", "/*", doubled quotes, backslash escapes, UTF-8
  void quoted(char q, uint32_t maxlen) {
    o.push_back(q);
    uint32_t n = r.below(maxlen + 1);
    for (uint32_t i = 0; i < n; i++) {
      uint32_t k = r.below(40);
      if (k == 0) { o.push_back(q); o.push_back(q); }
      else if (k == 1) { o.push_back('\\'); o.push_back(q); }
      else if (k == 2) o += "\\\\";
      else if (k == 3) o += "\\n";
      else if (k == 4) o += "\\\n";
      else if (k == 5) o += "\\u{767D}";
      else if (k == 6) o.push_back(q == '\'' ? '"' : '\'');
      else if (k == 7) o.push_back(';');
      else if (k == 8) o += "--";
      else if (k == 9) o += "/*";
      else if (k == 10) o += "*/";
      else if (k == 11) o.push_back('`');
      else if (k < 15) utf8_char();
      else {
        char c = (char)r.range(32, 126);
        if (c == q || c == '\\') c = ' ';
        o.push_back(c);
      }
    }
    o.push_back(q);
  }
  void backtick_ident() {
    o.push_back('`');
    uint32_t n = r.range(1, 20);
    for (uint32_t i = 0; i < n; i++) {
      uint32_t k = r.below(20);
      if (k == 0) utf8_char();
      else if (k == 1) o.push_back('\'');
      else if (k == 2) o.push_back(';');
      else {
        char c = (char)r.range(32, 126);
        if (c == '`') c = '_';
        o.push_back(c);
      }
    }
    o.push_back('`');
  }
  void comment_body(uint32_t maxlen, bool line) {
    uint32_t n = r.below(maxlen + 1);
    for (uint32_t i = 0; i < n; i++) {
      uint32_t k = r.below(30);
      if (k == 0) o.push_back('\'');
      else if (k == 1) o.push_back('"');
      else if (k == 2) o.push_back(';');
      else if (k == 3) o.push_back('`');
      else if (k == 4) utf8_char();
      else if (k == 5) o += line ? "/*" : "--";
      else if (k == 6) o.push_back('*');
      else {
        char c = (char)r.range(32, 126);
        if (!line && (c == '/' || c == '*')) c = ' ';
        o.push_back(c);
      }
    }
  }
  void block_comment() {
    uint32_t k = r.below(10);
    if (k == 0) { o += "/***/"; return; }
    o += "/*";
    comment_body(40, false);
    o += k == 1 ? "**/" : "*/";
  }
  void line_comment() {
    o += "--";
    uint32_t sp = r.below(3);
    for (uint32_t i = 0; i < sp; i++) o.push_back(' ');
    comment_body(40, true);
    o.push_back('\n');
  }
  void ws() {
    uint32_t k = r.below(20);
    if (k == 0) block_comment();
    else if (k == 1) line_comment();
    else if (k == 2) o += "\n  ";
    else if (k == 3) o += "\t";
    else o.push_back(' ');
  }
  void stress_value() {
    uint32_t k = r.below(10);
    if (k < 4) quoted('\'', 48);
    else if (k < 7) quoted('"', 48);
    else if (k < 8) backtick_ident();
    else if (k < 9) integer();
    else ident();
  }
  void stress_valid() {
    o += "SELECT";
    ws();
    uint32_t n = r.range(1, 4);
    for (uint32_t i = 0; i < n; i++) {
      if (i) { o += ","; ws(); }
      stress_value();
      if (r.chance(15)) { ws(); o += "AS"; ws(); backtick_ident(); }
    }
    ws();
    o += "FROM";
    ws();
    if (r.chance(50)) backtick_ident();
    else ident();
    if (r.chance(60)) {
      ws();
      o += "WHERE";
      ws();
      if (r.chance(50)) backtick_ident();
      else ident();
      ws();
      cmp_op();
      ws();
      stress_value();
    }
    if (r.chance(30)) { ws(); block_comment(); }
    if (r.chance(20)) { ws(); line_comment(); }
  }
  // one mutation per malformed statement, each tied to one error site of the reference
  void stress_malformed() {
    uint32_t k = r.below(38);
    auto head = [&]() { o += "SELECT "; stress_value(); o += ", "; };
    switch (k) {
      case 0: head(); o += "'unterminated "; ident(); break;                       // STR_EOF
      case 1: head(); o += "\"unterminated "; ident(); break;
      case 2: head(); o += "`unterminated "; ident(); break;                        // BT_EOF
      case 3: head(); o += "1 /* open "; ident(); break;                            // BC_EOF
      case 4: head(); o += "'raw\nnewline'"; break;                                 // STR_LF
      case 5: head(); o += "'raw\rreturn'"; break;                                  // STR_CR
      case 6: head(); o += "`raw\nnewline`"; break;                                 // BT_NL
      case 7: head(); o += "a ! b"; break;                                          // BANG
      case 8: head(); o += "1d"; break;                                             // NUM_INT
      case 9: head(); o += "0q"; break;                                             // NUM_ZERO
      case 10: head(); o += "1.5.2"; break;                                         // NUM_FLOAT
      case 11: head(); o += "12("; break;
      case 12: head(); o += "@1"; break;                                            // CFG_DIGIT
      case 13: head(); o += "@ "; break;                                            // CFG_EMPTY
      case 14: head(); o += "``"; break;                                            // BT_EMPTY
      case 15: head(); o += "$ "; break;                                            // QP_EMPTY
      case 16: head(); o += "$a"; break;                                            // QP_END
      case 17: head(); o += "$1x"; break;
      case 18: head(); o += "caf\xC3\xA9"; break;                                   // IDENT_END
      case 19: head(); o += "# 1"; break;                                           // INVALID_CHAR
      case 20: o += "SELECT a FROM t ORDER BY a ASC"; break;                        // quirk 1
      case 21: o += "SELECT $0"; break;                                             // quirk 2
      case 22: head(); o += "((1)"; break;                                          // unbalanced
      case 23: head(); o += "(1))"; break;
      case 24: head(); o += "[1, 2"; break;
      case 25: head(); o += "(1; 2)"; break;
      case 26: o += "INSERT INTO t (a, b) VALUES (1, 2), (3)"; break;               // ROW_WIDTH
      case 27: head(); o += "1234567890123456789012345678901234567890"; break;      // 40-digit integer
      case 28: o += "SELECT a FROM t LIMIT 18446744073709551616"; break;            // usize overflow
      case 29: o += "SELECT a FROM t LIMIT 0x"; break;                              // empty hex
      case 30: head(); o += "'\\u{110000}'"; break;                                 // InvalidEscapedUnicode
      case 31: head(); o += "- x"; break;                                           // quirk 5
      case 32: o += "CREATE TABLE t (a Int8) PRIMARY KEY a PRIMARY KEY a"; break;   // Conflicts
      case 33: o += "CREATE VIEW v AS SELECT 1"; break;                             // quirk 10
      case 34: o += "  "; break;                                                    // EmptyQuery
      case 35: o += "FROB x"; break;                                                // unrecognised
      case 36: head(); o += "a:b"; break;
      default: head(); o += "x NOT y"; break;                                       // NOT_INFIX list
    }
  }
  void config3() {
    if (r.chance(5)) stress_malformed();
    else stress_valid();
    o += ";\n";
  }

  // ---------------------------------------------------------------- config 4
  void atom() {
    uint32_t k = r.below(4);
    if (k == 0) integer();
    else if (k == 1) ident();
    else if (k == 2) sq_string(false);
    else { ident(); o.push_back('.'); ident(); }
  }
  void nested(uint32_t shape, uint32_t d) {
    static const char* const binops[] = {"OR", "XOR", "AND", "=", "<", "|", "^", "&", "<<", "+", "-", "*", "/", "%",
                                         ">=", "LIKE", "IN", ">>"};
    switch (shape) {
      case 0:
        for (uint32_t i = 0; i < d; i++) o.push_back('(');
        atom();
        for (uint32_t i = 0; i < d; i++) o.push_back(')');
        break;
      case 1:
        for (uint32_t i = 0; i < d; i++) o.push_back('[');
        atom();
        for (uint32_t i = 0; i < d; i++) o.push_back(']');
        break;
      case 2:
        for (uint32_t i = 0; i < d; i++) o += "{1:";
        atom();
        for (uint32_t i = 0; i < d; i++) o.push_back('}');
        break;
      case 3:
        for (uint32_t i = 0; i < d; i++) { ident(); o.push_back('('); }
        atom();
        for (uint32_t i = 0; i < d; i++) o.push_back(')');
        break;
      case 4:
        for (uint32_t i = 0; i < d; i++) { ident(); o.push_back('['); }
        atom();
        for (uint32_t i = 0; i < d; i++) o.push_back(']');
        break;
      case 5:
        for (uint32_t i = 0; i < d; i++) o += "(select ";
        atom();
        for (uint32_t i = 0; i < d; i++) o.push_back(')');
        break;
      case 6:
        for (uint32_t i = 0; i < d; i++) { o += "CASE WHEN "; ident(); o += " THEN "; }
        atom();
        for (uint32_t i = 0; i < d; i++) o += r.chance(50) ? " ELSE 0 END" : " END";
        break;
      case 7:  // left-deep chain (operators of random power)
        atom();
        for (uint32_t i = 0; i < d; i++) { o.push_back(' '); o += binops[r.below(18)]; o.push_back(' '); atom(); }
        break;
      case 8:  // right-deep: increasing power
        for (uint32_t i = 0; i < d; i++) { ident(); o.push_back(' '); o += binops[(i * 18) / d]; o.push_back(' '); }
        atom();
        break;
      case 9:
        for (uint32_t i = 0; i < d; i++) o += "NOT ";
        ident();
        break;
      case 10:
        for (uint32_t i = 0; i < d; i++) o.push_back('~');
        ident();
        break;
      case 11:
        ident();
        for (uint32_t i = 0; i < d; i++) { o += " BETWEEN "; integer(); o += " AND "; integer(); o += " AND "; ident(); }
        break;
      case 12:  // mixed brackets
        for (uint32_t i = 0; i < d; i++) o += (i % 3 == 0) ? "(" : (i % 3 == 1) ? "[" : "f(";
        atom();
        for (uint32_t i = d; i-- > 0;) o += (i % 3 == 1) ? "]" : ")";
        break;
      default:  // IF nesting
        for (uint32_t i = 0; i < d; i++) { o += "IF "; ident(); o += " THEN "; }
        atom();
        for (uint32_t i = 0; i < d; i++) o += " ELSE 0 END";
        break;
    }
  }
  void config4() {
    uint32_t d = 1u << r.below(9);  // 1..256
    uint32_t shape = r.below(14);
    if (r.chance(80)) {
      o += "SELECT ";
      nested(shape, d);
      o += " FROM ";
      ident();
      if (r.chance(30)) {
        o += " WHERE ";
        nested(r.below(14), 1u << r.below(4));
      }
    } else {
      o += "INSERT INTO ";
      ident();
      o += " VALUES (";
      nested(shape, d);
      o += ")";
    }
    o += ";\n";
  }

  void statement(int config) {
    if (config == 3) config3();
    else if (config == 4) config4();
    else config2();
  }
};

struct Block {
  std::string text;
  std::vector<uint32_t> ends;  // end offset of each statement within the block
};

void gen_block(int config, uint64_t seed, uint64_t k, Block& b) {
  b.text.clear();
  b.ends.clear();
  b.text.reserve(WL_BLOCK * 160);
  b.ends.reserve(WL_BLOCK);
  Gen g(seed * 0x9E3779B97F4A7C15ull + k * 0xD1B54A32D192ED03ull + 0x5EEDull, b.text);
  for (uint32_t i = 0; i < WL_BLOCK; i++) {
    g.statement(config);
    b.ends.push_back((uint32_t)b.text.size());
  }
}

struct Workload {
  std::vector<uint8_t> text;
  std::vector<uint64_t> offs;
};

}  // namespace

extern "C" {

// Generates statements until adding one more would exceed target_bytes (at least one statement).
void* nutdb_workload_create(int config, uint64_t seed, uint64_t target_bytes, int nthreads) {
  if (nthreads < 1) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  auto* w = new Workload();
  std::vector<Block> blocks;
  uint64_t have = 0;
  size_t done = 0;
  // estimate the number of blocks, generate them in parallel, extend if short
  while (have < target_bytes) {
    uint64_t per_block = done ? have / done : (uint64_t)WL_BLOCK * 100;
    size_t more = (size_t)((target_bytes - have) / (per_block ? per_block : 1) + 1);
    size_t first = blocks.size();
    blocks.resize(first + more);
    std::atomic<size_t> next{first};
    auto work = [&]() {
      for (;;) {
        size_t k = next.fetch_add(1);
        if (k >= blocks.size()) return;
        gen_block(config, seed, k, blocks[k]);
      }
    };
    std::vector<std::thread> th;
    int nt = (int)std::min<size_t>((size_t)nthreads, more);
    for (int i = 1; i < nt; i++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    for (size_t k = first; k < blocks.size(); k++) have += blocks[k].text.size();
    done = blocks.size();
  }
  // assemble: whole statements while they fit
  w->text.reserve(target_bytes + 64);
  w->offs.reserve(done * WL_BLOCK + 1);
  w->offs.push_back(0);
  bool full = false;
  for (size_t k = 0; k < blocks.size() && !full; k++) {
    const Block& b = blocks[k];
    uint32_t prev = 0;
    size_t base = w->text.size();
    uint32_t take = 0;
    for (uint32_t e : b.ends) {
      if (base + e > target_bytes && !(w->offs.size() == 1)) {
        full = true;
        break;
      }
      w->offs.push_back(base + e);
      take = e;
      prev = e;
    }
    (void)prev;
    w->text.insert(w->text.end(), b.text.begin(), b.text.begin() + take);
    blocks[k] = Block();
  }
  w->text.resize(w->text.size() + 64, 0);  // zero padding so 16-byte loads past the end are safe
  return w;
}
uint64_t nutdb_workload_bytes(void* h) { return ((Workload*)h)->offs.back(); }
uint64_t nutdb_workload_statements(void* h) { return ((Workload*)h)->offs.size() - 1; }
const uint8_t* nutdb_workload_text(void* h) { return ((Workload*)h)->text.data(); }
const uint64_t* nutdb_workload_offsets(void* h) { return ((Workload*)h)->offs.data(); }
void nutdb_workload_free(void* h) { delete (Workload*)h; }

}  // extern "C"
