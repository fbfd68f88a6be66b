// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not linked into libnutdb_gpu.so.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build or call anything in this directory.
//
// CPU restatement (C++17) of the reference tokenizer, following the Rust source function by
// function.  Citations are to /root/reference/src/parser/tokenizer/{utf8_iter,mod,token,error}.rs.
//
// PARITY STATUS: the Rust reference cannot be compiled in this image (no cargo/rustc), so this
// restatement is pinned against the reference's own unit-test vectors (tokenizer/mod.rs:576-782,
// utf8_iter.rs:284-307) restated in tests/test_oracle_lexer.py -- token level parity is PINNED by
// those vectors; everything they do not cover follows the source text only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <optional>

#include "../include/nutdb_gpu.h"

namespace ora {

// tokenizer/utf8_iter.rs:4-20
struct Span {
  size_t start = 0, end = 0;
  bool is_empty() const { return end == start; }
};

// tokenizer/utf8_iter.rs:22-38
struct Position {
  size_t line = 0, col = 0;
};

inline std::string encode_utf8(char32_t c) {
  std::string s;
  if (c < 0x80) {
    s.push_back((char)c);
  } else if (c < 0x800) {
    s.push_back((char)(0xC0 | (c >> 6)));
    s.push_back((char)(0x80 | (c & 0x3F)));
  } else if (c < 0x10000) {
    s.push_back((char)(0xE0 | (c >> 12)));
    s.push_back((char)(0x80 | ((c >> 6) & 0x3F)));
    s.push_back((char)(0x80 | (c & 0x3F)));
  } else {
    s.push_back((char)(0xF0 | (c >> 18)));
    s.push_back((char)(0x80 | ((c >> 12) & 0x3F)));
    s.push_back((char)(0x80 | ((c >> 6) & 0x3F)));
    s.push_back((char)(0x80 | (c & 0x3F)));
  }
  return s;
}

// tokenizer/utf8_iter.rs:40-237
class Utf8Iter {
 public:
  explicit Utf8Iter(std::string_view raw) : raw_(raw) {}
  size_t cursor() const { return cursor_; }
  size_t len() const { return raw_.size(); }
  std::string_view slice(const Span& s) const { return raw_.substr(s.start, s.end - s.start); }
  std::string_view slice_from_to(size_t a, size_t b) const { return raw_.substr(a, b - a); }
  std::string_view raw() const { return raw_; }
  void pin() { pinned_ = cursor_; }                                  // :80
  Span cut_from_pinned() const { return Span{pinned_, cursor_}; }    // :85

  // :89-116 -- rescans the prefix; '\r' (+ optional '\n') or '\n' => new line, '\t' => +4 cols,
  // any other char => +1 col.
  Position get_pos(size_t cursor) const {
    size_t col = 1, line = 1;
    size_t i = 0;
    while (i < cursor) {
      unsigned char x = (unsigned char)raw_[i];
      size_t n = x < 0x80 ? 1 : (x >= 0xF0 ? 4 : (x >= 0xE0 ? 3 : 2));
      if (x == '\r') {
        if (i + 1 < cursor && raw_[i + 1] == '\n') i += 1;  // next_if_eq(&'\n')
        line += 1;
        col = 1;
      } else if (x == '\n') {
        line += 1;
        col = 1;
      } else if (x == '\t') {
        col += 4;
      } else {
        col += 1;
      }
      i += n;
    }
    return Position{line, col};
  }
  Position get_current_pos() const { return get_pos(cursor_); }     // :118

  // :126-174
  std::optional<char32_t> peek() {
    if (peeked_len_ != 0) return peeked_;
    if (cursor_ == raw_.size()) return std::nullopt;
    peeked_len_ += 1;
    unsigned char x = (unsigned char)raw_[cursor_];
    if (x < 128) {
      peeked_ = x;
      return peeked_;
    }
    uint32_t init = x & (0x7F >> 2);
    peeked_len_ += 1;
    unsigned char y = (unsigned char)raw_[cursor_ + 1];
    uint32_t ch = (init << 6) | (y & 0x3F);
    if (x >= 0xE0) {
      peeked_len_ += 1;
      unsigned char z = (unsigned char)raw_[cursor_ + 2];
      uint32_t y_z = ((uint32_t)(y & 0x3F) << 6) | (z & 0x3F);
      ch = init << 12 | y_z;
      if (x >= 0xF0) {
        peeked_len_ += 1;
        unsigned char w = (unsigned char)raw_[cursor_ + 3];
        ch = (init & 7) << 18 | ((y_z << 6) | (w & 0x3F));
      }
    }
    peeked_ = ch;
    return peeked_;
  }
  std::optional<char32_t> next() {  // :177
    auto ch = peek();
    consume_peeked();
    return ch;
  }
  void consume_peeked() {  // :186
    cursor_ += peeked_len_;
    peeked_len_ = 0;
  }
  template <class P>
  Span take_while(P pred) {  // :192
    size_t start = cursor_;
    for (;;) {
      auto ch = peek();
      if (ch && pred(*ch)) {
        consume_peeked();
        continue;
      }
      break;
    }
    return Span{start, cursor_};
  }
  template <class P>
  void skip_while(P pred) {  // :210
    (void)take_while(pred);
  }
  void skip(char32_t c) {  // :226
    skip_while([c](char32_t x) { return x == c; });
  }

 private:
  std::string_view raw_;
  size_t cursor_ = 0;
  char32_t peeked_ = 0;
  uint8_t peeked_len_ = 0;
  size_t pinned_ = 0;
};

using TT = uint8_t;  // NUTDB_TT_*

// tokenizer/token.rs:89-115
struct Token {
  TT t = NUTDB_TT_EOF;
  Span span;
  bool is_whitespace() const { return t == NUTDB_TT_Whitespace || t == NUTDB_TT_Comment; }
  bool maybe_keyword() const { return t == NUTDB_TT_KeywordOrIdentifier; }
  bool is_terminator() const { return t == NUTDB_TT_EOF || t == NUTDB_TT_SemiColon; }
};

// tokenizer/error.rs:7-30
enum class TokErrType { UnexpectedEOF, UnexpectedChar, Incomplete };
struct TokenizeError {
  TokErrType t;
  std::string ctx;
  Position pos;
  size_t byte_pos;  // cursor the position was computed from (for record comparison)
  int site;         // NUTDB_LE_*
};

inline const char* tok_err_type_str(TokErrType t) {  // error.rs:14-22
  switch (t) {
    case TokErrType::UnexpectedEOF: return "Unexpected EOF";
    case TokErrType::UnexpectedChar: return "Unexpected Char";
    default: return "Incomplete Token";
  }
}

struct TokResult {
  bool ok;
  Token tok;
  TokenizeError err;
};

// tokenizer/mod.rs:12-543
class Tokenizer {
 public:
  explicit Tokenizer(std::string_view raw) : source(raw) {}
  Utf8Iter source;

  TokResult next_token() {  // :66-112
    source.pin();
    if (skip_whitespace()) return emit(NUTDB_TT_Whitespace);
    auto p = source.peek();
    if (!p) return emit(NUTDB_TT_EOF);
    char32_t ch = *p;
    switch (ch) {
      case '(': return consume_emit(NUTDB_TT_LParen);
      case ')': return consume_emit(NUTDB_TT_RParen);
      case '[': return consume_emit(NUTDB_TT_LBracket);
      case ']': return consume_emit(NUTDB_TT_RBracket);
      case '{': return consume_emit(NUTDB_TT_LBrace);
      case '}': return consume_emit(NUTDB_TT_RBrace);
      case ',': return consume_emit(NUTDB_TT_Comma);
      case ':': return consume_emit(NUTDB_TT_Colon);
      case '+': return consume_emit(NUTDB_TT_Plus);
      case '-': return tokenize_inline_comment_or_minus();
      case '*': return consume_emit(NUTDB_TT_Mul);
      case '/': return tokenize_block_comment_or_div();
      case '%': return consume_emit(NUTDB_TT_Mod);
      case '=': return consume_emit(NUTDB_TT_Eq);
      case '!': return tokenize_ne();
      case '<': return tokenize_lt();
      case '>': return tokenize_gt();
      case '&': return consume_emit(NUTDB_TT_BitAnd);
      case '|': return consume_emit(NUTDB_TT_BitOr);
      case '^': return consume_emit(NUTDB_TT_BitXor);
      case '~': return consume_emit(NUTDB_TT_BitNot);
      case ';': return consume_emit(NUTDB_TT_SemiColon);
      case '`': return tokenize_delimited_identifier();
      case '$': return tokenize_query_parameter();
      case '@': return tokenize_config_identifier();
      case '\'': return tokenize_string('\'', NUTDB_TT_EscapedSQStringLiteral);
      case '"': return tokenize_string('"', NUTDB_TT_EscapedDQStringLiteral);
      default: break;
    }
    if (is_ident_start(ch)) return tokenize_keyword_or_identifier();
    if (ch == '.' || is_digit(ch)) return tokenize_dot_or_numeric();
    return error(TokErrType::UnexpectedChar,
                 "'" + encode_utf8(ch) + "' is invalid outside string literal", NUTDB_LE_INVALID_CHAR);
  }

 private:
  static bool is_digit(char32_t c) { return c >= '0' && c <= '9'; }
  static bool is_ident_start(char32_t c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_';
  }
  static bool is_ident_char(char32_t c) { return is_ident_start(c) || is_digit(c); }
  static bool is_hex(char32_t c) {
    return is_digit(c) || (c >= 'A' && c <= 'F') || (c >= 'a' && c <= 'f');
  }

  TokResult emit(TT t) { return TokResult{true, Token{t, source.cut_from_pinned()}, {}}; }  // emit_token!(self, t)
  TokResult emit_on(TT t, Span s) { return TokResult{true, Token{t, s}, {}}; }             // emit_token!(t on span)
  TokResult consume_emit(TT t) {                                                            // :46-51
    source.consume_peeked();
    return emit(t);
  }
  TokResult error(TokErrType t, std::string ctx, int site) {  // emit_error! :33-41
    TokResult r;
    r.ok = false;
    r.err = TokenizeError{t, std::move(ctx), source.get_current_pos(), source.cursor(), site};
    return r;
  }

  // :115-184
  TokResult tokenize_string(char32_t quote, TT escaped_type) {
    source.consume_peeked();
    bool escaped = false;
    source.pin();
    for (;;) {
      auto p = source.peek();
      if (!p) return error(TokErrType::UnexpectedEOF, "string literal is not complete", NUTDB_LE_STR_EOF);
      char32_t ch = *p;
      if (ch == quote) {
        Span span = source.cut_from_pinned();
        source.consume_peeked();
        auto n = source.peek();
        if (n && *n == quote) {
          source.consume_peeked();
          escaped = true;
        } else {
          return escaped ? emit_on(escaped_type, span) : emit_on(NUTDB_TT_RawStringLiteral, span);
        }
      } else if (ch == '\\') {
        source.consume_peeked();
        auto next_ch = source.next();
        if (next_ch && *next_ch == '\r') {
          auto n = source.peek();
          if (n && *n == '\n') source.consume_peeked();
        }
        escaped = true;
      } else if (ch == '\r') {
        return error(TokErrType::UnexpectedChar,
                     "\\r in string is supported but should be escaped by '\\'", NUTDB_LE_STR_CR);
      } else if (ch == '\n') {
        return error(TokErrType::UnexpectedChar,
                     "\\n in string is supported but should be escaped by '\\'", NUTDB_LE_STR_LF);
      } else {
        source.consume_peeked();
      }
    }
  }

  // :191-260
  TokResult tokenize_dot_or_numeric() {
    source.pin();
    Span span = source.take_while(is_digit);
    if (source.slice(span) == "0") {
      auto p = source.peek();
      if (p && (*p == 'x' || *p == 'X')) {
        source.consume_peeked();
        Span hs = source.take_while(is_hex);
        return emit_on(NUTDB_TT_HexLiteral, hs);
      } else if (p && *p == '.') {
        // is float
      } else {
        if (auto bad = invalid_end_of_numeric(p))
          return error(TokErrType::UnexpectedChar,
                       "'" + encode_utf8(*bad) + "' is invalid in numeric literal", NUTDB_LE_NUM_ZERO);
        return emit_on(NUTDB_TT_IntegerLiteral, span);
      }
    }
    {
      auto p = source.peek();
      if (p && *p == '.') {
        source.consume_peeked();
      } else {
        if (auto bad = invalid_end_of_numeric(p))
          return error(TokErrType::UnexpectedChar,
                       "'" + encode_utf8(*bad) + "' cannot be a part of integer literal", NUTDB_LE_NUM_INT);
        return emit_on(NUTDB_TT_IntegerLiteral, span);
      }
    }
    source.skip_while(is_digit);
    Span full = source.cut_from_pinned();
    if (source.slice(full) == ".") return emit_on(NUTDB_TT_Dot, full);
    if (auto bad = invalid_end_of_numeric(source.peek()))
      return error(TokErrType::UnexpectedChar,
                   "'" + encode_utf8(*bad) + "' cannot be a part of float literal", NUTDB_LE_NUM_FLOAT);
    return emit_on(NUTDB_TT_FloatLiteral, full);
  }

  // :262-282
  TokResult tokenize_keyword_or_identifier() {
    Span span = source.take_while(is_ident_char);
    if (auto bad = invalid_end_of_identifier(source.peek()))
      return error(TokErrType::UnexpectedChar,
                   "'" + encode_utf8(*bad) + "' cannot be a part of identifier or keyword", NUTDB_LE_IDENT_END);
    return emit_on(NUTDB_TT_KeywordOrIdentifier, span);
  }

  // :284-311
  TokResult tokenize_config_identifier() {
    source.consume_peeked();
    auto p = source.peek();
    if (p && is_digit(*p))
      return error(TokErrType::UnexpectedChar, "config identifier cannot starts with numbers", NUTDB_LE_CFG_DIGIT);
    Span span = source.take_while(is_ident_char);
    if (auto bad = invalid_end_of_identifier(source.peek()))
      return error(TokErrType::UnexpectedChar,
                   "'" + encode_utf8(*bad) + "' cannot be a part of config identifier", NUTDB_LE_CFG_END);
    if (span.is_empty()) return error(TokErrType::Incomplete, "identifier should have name", NUTDB_LE_CFG_EMPTY);
    return emit_on(NUTDB_TT_ConfigIdentifier, span);
  }

  // :313-345
  TokResult tokenize_delimited_identifier() {
    source.consume_peeked();
    Span span = source.take_while([](char32_t c) { return !(c == '`' || c == '\r' || c == '\n'); });
    if (span.is_empty())
      return error(TokErrType::Incomplete, "delimited identifier cannot be an empty string", NUTDB_LE_BT_EMPTY);
    auto p = source.peek();
    if (p && *p == '`') {
      source.consume_peeked();
      return emit_on(NUTDB_TT_DelimitedIdentifier, span);
    }
    if (p)
      return error(TokErrType::UnexpectedChar, "'\\r' or '\\n' cannot be a part of delimited identifier",
                   NUTDB_LE_BT_NL);
    return error(TokErrType::UnexpectedEOF, "delimited identifier is not complete", NUTDB_LE_BT_EOF);
  }

  // :347-365
  TokResult tokenize_query_parameter() {
    source.consume_peeked();
    Span span = source.take_while(is_digit);
    if (auto bad = invalid_end_of_query_parameter(source.peek()))
      return error(TokErrType::UnexpectedChar,
                   "'" + encode_utf8(*bad) + "' cannot be a part of query parameter", NUTDB_LE_QP_END);
    if (span.is_empty())
      return error(TokErrType::Incomplete, "query parameter should have an index", NUTDB_LE_QP_EMPTY);
    return emit_on(NUTDB_TT_QueryParameter, span);
  }

  // :367-378
  TokResult tokenize_inline_comment_or_minus() {
    source.consume_peeked();
    auto p = source.peek();
    if (p && *p == '-') {
      source.consume_peeked();
      return tokenize_inline_comment_body();
    }
    return emit(NUTDB_TT_Minus);
  }
  // :380-391
  TokResult tokenize_block_comment_or_div() {
    source.consume_peeked();
    auto p = source.peek();
    if (p && *p == '*') {
      source.consume_peeked();
      return tokenize_block_comment_body();
    }
    return emit(NUTDB_TT_Div);
  }
  // :393-403
  TokResult tokenize_ne() {
    source.consume_peeked();
    auto p = source.peek();
    if (p && *p == '=') {
      source.consume_peeked();
      return emit(NUTDB_TT_NotEq);
    }
    return error(TokErrType::UnexpectedChar, "'!' can only be used with '='", NUTDB_LE_BANG);
  }
  // :405-416
  TokResult tokenize_lt() {
    source.consume_peeked();
    auto p = source.peek();
    if (p && *p == '=') return consume_emit(NUTDB_TT_LtEq);
    if (p && *p == '>') return consume_emit(NUTDB_TT_NotEq);
    if (p && *p == '<') return consume_emit(NUTDB_TT_BitLShift);
    return emit(NUTDB_TT_Lt);
  }
  // :418-428
  TokResult tokenize_gt() {
    source.consume_peeked();
    auto p = source.peek();
    if (p && *p == '=') return consume_emit(NUTDB_TT_GtEq);
    if (p && *p == '>') return consume_emit(NUTDB_TT_BitRShift);
    return emit(NUTDB_TT_Gt);
  }
  // :430-437
  TokResult tokenize_inline_comment_body() {
    source.skip(' ');
    Span span = source.take_while([](char32_t c) { return c != '\n' && c != '\r'; });
    return emit_on(NUTDB_TT_Comment, span);
  }
  // :439-468
  TokResult tokenize_block_comment_body() {
    size_t start = source.cursor();
    size_t end = source.cursor();
    uint8_t comment_end = 0;
    for (;;) {
      if (comment_end == 2) break;
      auto p = source.peek();
      if (!p) return error(TokErrType::UnexpectedEOF, "block comment is not complete", NUTDB_LE_BC_EOF);
      char32_t ch = *p;
      source.consume_peeked();
      if (comment_end == 1 && ch == '/') {
        comment_end = 2;
      } else {
        comment_end = (ch == '*') ? 1 : 0;
      }
      if (comment_end == 0) end = source.cursor();
    }
    return emit_on(NUTDB_TT_Comment, Span{start, end});
  }

  // :473-477
  bool skip_whitespace() {
    size_t start = source.cursor();
    source.skip_while([](char32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; });
    return start != source.cursor();
  }

  // :486-503
  static std::optional<char32_t> invalid_end_of_identifier(std::optional<char32_t> ch) {
    if (!ch) return std::nullopt;
    switch (*ch) {
      case '+': case '-': case '*': case '/': case '%': case '&': case '|': case '^': case '>': case '<': case '=':
      case '!':
      case '.': case ',': case ';':
      case '[': case ']': case '(': case ')': case '{': case '}':
      case '\t': case '\n': case '\r': case ' ':
        return std::nullopt;
      default: return ch;
    }
  }
  // :506-523
  static std::optional<char32_t> invalid_end_of_query_parameter(std::optional<char32_t> ch) {
    if (!ch) return std::nullopt;
    switch (*ch) {
      case '+': case '-': case '*': case '/': case '%': case '&': case '|': case '^': case '>': case '<': case '=':
      case '!':
      case ',': case ':': case ';':
      case ']': case ')': case '}':
      case '\t': case '\n': case '\r': case ' ':
        return std::nullopt;
      default: return ch;
    }
  }
  // :526-543
  static std::optional<char32_t> invalid_end_of_numeric(std::optional<char32_t> ch) {
    return invalid_end_of_query_parameter(ch);  // identical character set (:529-540 vs :509-520)
  }
};

}  // namespace ora
