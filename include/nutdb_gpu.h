/*
 * nutdb_gpu.h -- C ABI of the B200-native SQL lexer/parser (libnutdb_gpu.so).
 *
 * Drop-in boundary for the ONE hot path of nutdb/nutdb: `nutdb::parser::Parser::parse`
 * (reference: src/parser/mod.rs:26-29, exported through src/lib.rs:3-4).  The reference has
 * no FFI of its own; a maintainer binds these entry points from a `-sys` crate
 *  and re-hydrates the flat arrays into `nutdb::parser::Statement`.
 *
 * Semantics: statement i of a batch is the byte range sql[stmt_off[i] .. stmt_off[i+1]) and is
 * parsed EXACTLY as `Parser::parse(&sql[stmt_off[i]..stmt_off[i+1]])` would be
 * (src/parser/mod.rs:128-180): one statement, parsing stops at the first `;`/EOF in
 * statement-final position, first error wins, all spans and (line,col) positions are
 * relative to the statement's own first byte.  Input must be valid UTF-8 (the `&str`
 * precondition of the reference).
 *
 * There is no CPU fallback: every entry point that computes needs a CUDA device and
 * returns NUTDB_E_CUDA otherwise.
 *
 * This header also fixes the numeric codes of the flat output format (token types, node
 * kinds, error codes).  The CPU oracle under oracle/ includes it for those constants only.
 */
#ifndef NUTDB_GPU_H
#define NUTDB_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Token types.  Ordinals 0..39 are the declaration order of `enum TokenType`
 * (reference: src/parser/tokenizer/token.rs:5-87).  40 is ours: a lex failure in token
 * position ("poison"), raised as ParseError::LexError only if the parser pulls it
 * (reference behaviour: lazy lexing, src/parser/mod.rs:1853-1883).
 * ---------------------------------------------------------------------------------------- */
enum {
  NUTDB_TT_KeywordOrIdentifier = 0,
  NUTDB_TT_DelimitedIdentifier = 1,
  NUTDB_TT_ConfigIdentifier = 2,
  NUTDB_TT_QueryParameter = 3,
  NUTDB_TT_RawStringLiteral = 4,
  NUTDB_TT_EscapedSQStringLiteral = 5,
  NUTDB_TT_EscapedDQStringLiteral = 6,
  NUTDB_TT_IntegerLiteral = 7,
  NUTDB_TT_FloatLiteral = 8,
  NUTDB_TT_HexLiteral = 9,
  NUTDB_TT_Comma = 10,
  NUTDB_TT_Dot = 11,
  NUTDB_TT_Colon = 12,
  NUTDB_TT_SemiColon = 13,
  NUTDB_TT_Plus = 14,
  NUTDB_TT_Minus = 15,
  NUTDB_TT_Mul = 16,
  NUTDB_TT_Div = 17,
  NUTDB_TT_Mod = 18,
  NUTDB_TT_Eq = 19,
  NUTDB_TT_NotEq = 20,
  NUTDB_TT_Lt = 21,
  NUTDB_TT_Gt = 22,
  NUTDB_TT_LtEq = 23,
  NUTDB_TT_GtEq = 24,
  NUTDB_TT_LParen = 25,
  NUTDB_TT_RParen = 26,
  NUTDB_TT_LBracket = 27,
  NUTDB_TT_RBracket = 28,
  NUTDB_TT_LBrace = 29,
  NUTDB_TT_RBrace = 30,
  NUTDB_TT_BitAnd = 31,
  NUTDB_TT_BitOr = 32,
  NUTDB_TT_BitXor = 33,
  NUTDB_TT_BitNot = 34,
  NUTDB_TT_BitLShift = 35,
  NUTDB_TT_BitRShift = 36,
  NUTDB_TT_Comment = 37,
  NUTDB_TT_Whitespace = 38,
  NUTDB_TT_EOF = 39,
  NUTDB_TT_POISON = 40, /* start = byte offset of the error, end = NUTDB_LE_* code */
  NUTDB_TT_COUNT = 41
};

/* Lex error sites (reference: src/parser/tokenizer/mod.rs, line cited per code).
 * The TokenizeErrorType of each is fixed by its site (tokenizer/error.rs:7-12). */
enum {
  NUTDB_LE_INVALID_CHAR = 1, /* :105 UnexpectedChar "'{c}' is invalid outside string literal" */
  NUTDB_LE_STR_CR = 2,       /* :161 UnexpectedChar "\r in string is supported but ..." */
  NUTDB_LE_STR_LF = 3,       /* :168 UnexpectedChar "\n in string is supported but ..." */
  NUTDB_LE_STR_EOF = 4,      /* :178 UnexpectedEOF  "string literal is not complete" */
  NUTDB_LE_NUM_ZERO = 5,     /* :216 UnexpectedChar "'{c}' is invalid in numeric literal" */
  NUTDB_LE_NUM_INT = 6,      /* :232 UnexpectedChar "'{c}' cannot be a part of integer literal" */
  NUTDB_LE_NUM_FLOAT = 7,    /* :252 UnexpectedChar "'{c}' cannot be a part of float literal" */
  NUTDB_LE_IDENT_END = 8,    /* :273 UnexpectedChar "'{c}' cannot be a part of identifier or keyword" */
  NUTDB_LE_CFG_DIGIT = 9,    /* :291 UnexpectedChar "config identifier cannot starts with numbers" */
  NUTDB_LE_CFG_END = 10,     /* :301 UnexpectedChar "'{c}' cannot be a part of config identifier" */
  NUTDB_LE_CFG_EMPTY = 11,   /* :307 Incomplete     "identifier should have name" */
  NUTDB_LE_BT_EMPTY = 12,    /* :324 Incomplete     "delimited identifier cannot be an empty string" */
  NUTDB_LE_BT_NL = 13,       /* :336 UnexpectedChar "'\r' or '\n' cannot be a part of delimited identifier" */
  NUTDB_LE_BT_EOF = 14,      /* :342 UnexpectedEOF  "delimited identifier is not complete" */
  NUTDB_LE_QP_END = 15,      /* :355 UnexpectedChar "'{c}' cannot be a part of query parameter" */
  NUTDB_LE_QP_EMPTY = 16,    /* :361 Incomplete     "query parameter should have an index" */
  NUTDB_LE_BANG = 17,        /* :401 UnexpectedChar "'!' can only be used with '='" */
  NUTDB_LE_BC_EOF = 18,      /* :463 UnexpectedEOF  "block comment is not complete" */
  NUTDB_LE_COUNT = 19
};

/* Keyword ids: 1 + index in declaration order of src/parser/keyword.rs:13-148 (115 words).
 * 0 = not a keyword.  Produced by the lexer's perfect hash as a side array; the token type
 * stays KeywordOrIdentifier because keywords are contextual in the reference. */
#define NUTDB_KW_COUNT 115

/* ------------------------------------------------------------------------------------------
 * Flat AST.  Nodes of one statement are stored in POST-ORDER (children before parent, in
 * source order), so every subtree is a contiguous index range [subtree_start, root].
 * All indices are relative to the statement's node_begin; spans are byte offsets relative
 * to the statement's first byte.
 *   leaf     : a,b = payload (usually the byte span [a,b) of the source slice)
 *   interior : a = subtree_start (index of its left-most descendant; == own index if it has
 *              no children), b = number of direct children
 *   parent   : index of the parent node, 0xFFFFFFFF for the root (= last node).
 * An NK_IDENT leaf whose aux bit0 is set owns the NK_QUAL leaf directly before it
 * (`qualifier.name`, reference: src/parser/ast/item.rs:78-81).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint8_t kind;    /* NUTDB_NK_* */
  uint8_t sub;     /* operator / variant, see each kind */
  uint16_t aux;    /* flag bits, see each kind */
  uint32_t parent; /* 0xFFFFFFFF for the root */
  uint32_t a;
  uint32_t b;
} NutdbNode;

/* The WIRE form of a node: ONE 32-bit word -- what the device produces and what crosses PCIe (NutdbBatch.pnode).  Same
 * order (post-order per statement), same kind / sub.  It carries everything NutdbNode does: a re-hydrator walks a
 * statement's words front to back with a running byte position `pos` (0 at the statement's start) and a stack, and
 * never needs `parent`:
 *   bits 0-6 kind, 7-11 sub, 12 = bit 0 of aux, then
 *   interior : bits 13-31 = subtree SIZE: the subtree is the `size` nodes in front of this one, so subtree_start =
 *              index - size.  NUTDB_PN_SIZE_EXT: see the side table.
 *   leaf     : bits 13-22 = gap, bits 23-31 = length: the span is [pos + gap, pos + gap + length) and pos moves to its
 *              end (leaves come in text order).  length == NUTDB_PN_LEN_SPECIAL marks the two special forms:
 *              gap == NUTDB_PN_GAP_NOSPAN: the leaf has no source span (a = b = 0; pos does not move);
 *              gap == NUTDB_PN_GAP_EXT: see the side table.
 * Side table (NutdbBatch.ext, sorted by batch-global node index; rare: a literal of 511 bytes or more, a kilobyte of
 * comment in front of a token, a subtree of half a million nodes): the exact fields of a node that does not fit --
 * hdr = kind | sub << 8 | aux << 16; leaf: a = span start, b = span length (pos moves to a + b); interior: a = subtree
 * start (statement relative).
 * nutdb_batch_expand_nodes() turns a batch's wire nodes into NutdbNode records on the host. */
#define NUTDB_PN_SUB_SHIFT 7
#define NUTDB_PN_FLAG_SHIFT 12
#define NUTDB_PN_SIZE_SHIFT 13
#define NUTDB_PN_SIZE_MAX 0x7FFFEu
#define NUTDB_PN_SIZE_EXT 0x7FFFFu
#define NUTDB_PN_GAP_SHIFT 13
#define NUTDB_PN_GAP_MAX 1021u
#define NUTDB_PN_GAP_EXT 1022u
#define NUTDB_PN_GAP_NOSPAN 1023u
#define NUTDB_PN_LEN_SHIFT 23
#define NUTDB_PN_LEN_MAX 510u
#define NUTDB_PN_LEN_SPECIAL 511u
typedef struct {
  uint32_t index; /* node index in the batch */
  uint32_t hdr;   /* kind | sub << 8 | aux << 16 */
  uint32_t a, b;
} NutdbNodeExt;

#define NUTDB_NO_PARENT 0xFFFFFFFFu
#define NUTDB_NK_FIRST_INTERIOR 32

enum {
  /* ---- leaves ---- */
  NUTDB_NK_NAME = 1,        /* an identifier string (must_parse_identifier_string, mod.rs:1682); span */
  NUTDB_NK_ALIAS = 2,       /* `AS name` applying to the previous sibling (mod.rs:563-578); span */
  NUTDB_NK_QUAL = 3,        /* qualifier of the NK_IDENT that follows; span */
  NUTDB_NK_IDENT = 4,       /* Expr::Identifier; sub 0=Word 1=Wildcard; aux bit0=has qualifier; span */
  NUTDB_NK_QPARAM = 5,      /* Expr::QueryParameter; span of the index integer; aux bit0=hex (mod.rs:1311) */
  NUTDB_NK_LIT_INT = 6,     /* Literal::Integer; sub bit0=negative; aux bit0=hex; span of digits */
  NUTDB_NK_LIT_FLOAT = 7,   /* Literal::Float; sub bit0=negated; span of lexeme */
  NUTDB_NK_LIT_STR = 8,     /* Literal::String; sub 0=raw 1=single-quote-escaped 2=double-quote-escaped; span of content */
  NUTDB_NK_LIT_BOOL = 9,    /* Literal::Boolean; sub=value; a=b=0 (may come from constant folding) */
  NUTDB_NK_LIT_NULL = 10,   /* Literal::Null; a=b=0 */
  NUTDB_NK_LIT_INTERVAL = 11, /* Literal::Interval; sub=unit 0..5 (Second..Year); aux bit0=hex; span of integer */
  NUTDB_NK_NUM = 12,        /* integer parameter (must_parse_integer_literal, mod.rs:1815); aux bit0=hex; span */
  NUTDB_NK_STR = 13,        /* string parameter (must_parse_string_literal, mod.rs:1833); sub as LIT_STR; span */
  NUTDB_NK_DT_SCALAR = 14,  /* ScalarDataType without parameter; sub = index 0..25 in ast/item.rs:15-50 order */
  NUTDB_NK_ORDER_DESC = 15, /* OrderDirection::DESC for the previous key */
  NUTDB_NK_FN_NAME = 16,    /* FnName::Others(name): first child of NK_FNCALL; span */
  NUTDB_NK_ENT_NAME = 17,   /* alter entity; sub 0=Column 1=Index 2=Constraint 3=Table(no span); span */
  NUTDB_NK_POS_FIRST = 18,  /* EntityPosition::First */
  NUTDB_NK_POS_AFTER = 19,  /* EntityPosition::After(name); span */
  NUTDB_NK_STRATEGY = 20,   /* ViewDefinition.strategy (`UPDATE BY name`); span */
  /* ---- interior ---- */
  NUTDB_NK_STMT_SELECT = 32,
  NUTDB_NK_STMT_INSERT = 33,   /* children: NAME(table) NAME(col)* data(ROWS|query|FNCALL) */
  NUTDB_NK_ROWS = 34,          /* children: ROW+ ; column_size = children of first ROW */
  NUTDB_NK_ROW = 35,
  NUTDB_NK_STMT_EXPLAIN = 36,
  NUTDB_NK_STMT_ALTER = 37,    /* sub 0=Add 1=Drop 2=Rename; aux bit0=if_[not_]exists; children: NAME(table) ... */
  NUTDB_NK_STMT_CREATE = 38,   /* aux bit0=if_not_exists; child TABLEDEF|VIEWDEF */
  NUTDB_NK_TABLEDEF = 39,      /* children: NAME, (COLDEF|INDEXDEF|CONSTRDEF)+, (ATTR_PK|ATTR_ORDER|ATTR_PART|STR)* in source order */
  NUTDB_NK_VIEWDEF = 40,       /* children: NAME, (STRATEGY|ATTR_PK|ATTR_ORDER|ATTR_PART|STR)*, query */
  NUTDB_NK_COLDEF = 41,        /* children: NAME, datatype, (ATTR_DEFAULT|STR)* */
  NUTDB_NK_INDEXDEF = 42,      /* children: NAME, FNCALL */
  NUTDB_NK_CONSTRDEF = 43,     /* children: NAME, expr */
  NUTDB_NK_ATTR_PK = 44,
  NUTDB_NK_ATTR_ORDER = 45,
  NUTDB_NK_ATTR_PART = 46,
  NUTDB_NK_ATTR_DEFAULT = 47,
  NUTDB_NK_STMT_DESCRIBE = 48, /* sub 0=Table 1=View 2=Database; child NAME unless Database */
  NUTDB_NK_STMT_DROP = 49,     /* sub 0=Table 1=View; aux bit0=if_exists; child NAME */
  NUTDB_NK_STMT_TRUNCATE = 50, /* same */
  NUTDB_NK_STMT_OPTIMIZE = 51, /* children NAME [expr] */
  NUTDB_NK_STMT_SET = 52,      /* children NAME(config) expr */
  NUTDB_NK_QUERY_BODY = 53,    /* children: WITH? DISTINCT? COLS FROM? JOIN* WHERE? GROUPBY? HAVING? ORDERBY? LIMIT? */
  NUTDB_NK_QUERY_UNION = 54,   /* sub 0=UnionAll 1=UnionDistinct 2=Intersect 3=Except; children left,right */
  NUTDB_NK_WITH = 55,          /* children: (NAME query)+ */
  NUTDB_NK_DISTINCT = 56,      /* aux bit0 = has ON(...); children (expr ALIAS?)* */
  NUTDB_NK_COLS = 57,          /* children (expr ALIAS?)+ */
  NUTDB_NK_FROM = 58,          /* children source ALIAS? */
  NUTDB_NK_JOIN = 59,          /* sub JoinType 0..8 (ast/query.rs:100-111); aux bit0 = USING; children source ALIAS? (expr | IDENT+) */
  NUTDB_NK_WHERE = 60,
  NUTDB_NK_GROUPBY = 61,
  NUTDB_NK_HAVING = 62,
  NUTDB_NK_ORDERBY = 63,       /* children (expr ALIAS? ORDER_DESC?)+ */
  NUTDB_NK_LIMIT = 64,         /* sub 0=`n` 1=`o, n` 2=`n OFFSET o`; aux bit0=with_ties; children NUM NUM? in source order */
  NUTDB_NK_COLLECTION = 65,    /* sub 0=Tuple 1=Map 2=Array */
  NUTDB_NK_UNARY = 66,         /* sub UnaryOperator 0..3 (ast/item.rs:128-134) */
  NUTDB_NK_BINARY = 67,        /* sub BinaryOperator 0..25 (ast/item.rs:136-164) */
  NUTDB_NK_FNCALL = 68,        /* sub FnName 0..7 (ast/item.rs:166-181); if 7 (Others) first child is FN_NAME */
  NUTDB_NK_DT_PARAM = 69,      /* sub scalar index (16,17,21,22); child NUM */
  NUTDB_NK_DT_COMPOUND = 70    /* sub 0=Array 1=Enum 2=Tuple 3=Map 4=Dictionary 5=Nullable; Enum children (STR NUM?)+; Map children key,value in SOURCE order */
};

/* ------------------------------------------------------------------------------------------
 * Per-statement result and error records.
 * ---------------------------------------------------------------------------------------- */
enum {
  NUTDB_ST_OK = 0,
  NUTDB_ST_LEX_ERROR = 1,    /* ParseError::LexError  (src/parser/error.rs:10) */
  NUTDB_ST_SYNTAX_ERROR = 2, /* ParseError::SyntaxError (src/parser/error.rs:12) */
  NUTDB_ST_LIMIT = 3,        /* nesting deeper than the device parser's stack (no reference equivalent; the reference would overflow its call stack) */
  NUTDB_ST_REFERENCE_PANIC = 4 /* the reference panics instead of returning: an escaped string literal that ends in a lone
                                  backslash after `\u` swallowed its partner reaches unreachable!() (literal.rs:63) */
};

/* SyntaxError variants (src/parser/error.rs:17-56), in declaration order, 1-based. */
enum {
  NUTDB_SE_NotExpectedTokenTypes = 1, /* a = expected list id (NUTDB_EL_*), b = actual token type */
  NUTDB_SE_NotExpectedKeywords = 2,   /* a = keyword list id (NUTDB_KL_* or 1000+keyword id), [b,c) = span of actual word */
  NUTDB_SE_ParseFail = 3,             /* a = message id (NUTDB_PF_*) */
  NUTDB_SE_EmptyQuery = 4,
  NUTDB_SE_InvalidEscapedUnicode = 5, /* [b,c) = span of the hex digits */
  NUTDB_SE_InvalidFloatLiteral = 6,   /* unreachable for lexer-produced floats */
  NUTDB_SE_InvalidHexLiteral = 7,     /* [b,c) = span of raw digits */
  NUTDB_SE_InvalidIntegerLiteral = 8, /* [b,c) = span of raw digits */
  NUTDB_SE_Conflicts = 9              /* a = NUTDB_CF_*; for ROW_WIDTH b = this row, c = previous rows */
};

/* expected-token-type lists of the next_expect! sites (src/parser/mod.rs:68-91) */
enum {
  NUTDB_EL_RParen = 1,
  NUTDB_EL_LParen = 2,
  NUTDB_EL_RBracket = 3,
  NUTDB_EL_RBrace = 4,
  NUTDB_EL_NegLiteral = 5,   /* [IntegerLiteral, HexLiteral, FloatLiteral]          mod.rs:1260 */
  NUTDB_EL_Identifier = 6,   /* [DelimitedIdentifier, KeywordOrIdentifier, Mul]     mod.rs:1512,1526 */
  NUTDB_EL_Colon = 7,
  NUTDB_EL_Keyword = 8,      /* [KeywordOrIdentifier]                               mod.rs:1637,1651,1667 */
  NUTDB_EL_IdentString = 9,  /* [KeywordOrIdentifier, DelimitedIdentifier]          mod.rs:1683 */
  NUTDB_EL_IntLiteral = 10,  /* [IntegerLiteral, HexLiteral]                        mod.rs:1816 */
  NUTDB_EL_StrLiteral = 11,  /* [Raw, EscapedSQ, EscapedDQ]                         mod.rs:1834 */
  NUTDB_EL_ConfigIdent = 12, /* mod.rs:1185 */
  NUTDB_EL_Eq = 13,          /* mod.rs:1189 */
  NUTDB_EL_Comma = 14,       /* mod.rs:1777 */
  NUTDB_EL_Prefix = 15,      /* the 16-type list at mod.rs:1314-1339 */
  NUTDB_EL_COUNT = 16
};

/* expected-keyword lists of the must_parse_one_of_keywords sites; single-keyword sites use
 * 1000 + keyword id. */
enum {
  NUTDB_KL_WITH_SELECT = 1,      /* mod.rs:221 */
  NUTDB_KL_ALL_DISTINCT = 2,     /* :259 */
  NUTDB_KL_ON_USING = 3,         /* :419 */
  NUTDB_KL_INSERT_SOURCE = 4,    /* :609 [values, from, select, with] */
  NUTDB_KL_TABLE_VIEW = 5,       /* :697,:1092,:1123 */
  NUTDB_KL_TABLE_ATTRS = 6,      /* :752 [primary, order, partition, comment] */
  NUTDB_KL_VIEW_ATTRS = 7,       /* :821 [as, update, primary, order, partition, comment] */
  NUTDB_KL_COLUMN_ATTRS = 8,     /* :946 [default, comment] */
  NUTDB_KL_ALTER_ACTION = 9,     /* :987 [add, drop, rename] */
  NUTDB_KL_ADD_ENTITY = 10,      /* :995 [column, index, constraint] */
  NUTDB_KL_DROP_ENTITY = 11,     /* :1025 [column, index, constraint, partition] */
  NUTDB_KL_RENAME_ENTITY = 12,   /* :1041 [column, index, constraint, table] */
  NUTDB_KL_DESCRIBE_ENTITY = 13, /* :1071 [table, view, database] */
  NUTDB_KL_NOT_INFIX = 14,       /* :1402 [in, like, ilike, between, exists] */
  NUTDB_KL_NOT_NULL = 15,        /* :1431 [not, null] */
  NUTDB_KL_INTERVAL_UNIT = 16,   /* :1493 */
  NUTDB_KL_DATATYPE = 17,        /* :1689-1703 (32 words) */
  NUTDB_KL_CASE_NEXT = 18,       /* :1601 [when, else, end] */
  NUTDB_KL_VIEW_NEEDS_UPDATE = 19, /* :826 expected [update], actual is the literal "as" */
  NUTDB_KL_COUNT = 20
};
#define NUTDB_KL_SINGLE 1000

enum {
  NUTDB_PF_START_KEYWORD = 1,  /* mod.rs:146 "statements should start with a keyword" */
  NUTDB_PF_MORE_THAN_ONE = 2,  /* :169 "more than one statement" */
  NUTDB_PF_UNRECOGNIZED = 3,   /* :175 "cannot recognize statement" */
  NUTDB_PF_NOT_SUBQUERY = 4,   /* :337 "not a subquery" */
  NUTDB_PF_QUERY_SOURCE = 5,   /* :557 "query source must be a subquery, a table function or a table" */
  NUTDB_PF_INSERT_SOURCE = 6,  /* :616 "insert source must be a subquery, values, or a function call" */
  NUTDB_PF_INDEXER = 7,        /* :926 "indexer must be a function call" */
  NUTDB_PF_NOT_EXISTS_ARGS = 8,/* :1417 "`not exists` should have arguments" */
  NUTDB_PF_EXISTS_ARGS = 9     /* :1454 (unreachable, SURVEY App.B quirk 7) */
};

enum {
  NUTDB_CF_PRIMARY_KEY = 1,  /* mod.rs:755,:847 */
  NUTDB_CF_ORDER_BY = 2,     /* :767,:859 */
  NUTDB_CF_PARTITION_BY = 3, /* :779,:872 */
  NUTDB_CF_COMMENT = 4,      /* :791,:884,:959 */
  NUTDB_CF_UPDATE_BY = 5,    /* :836 */
  NUTDB_CF_DEFAULT = 6,      /* :949 */
  NUTDB_CF_ROW_WIDTH = 7     /* :659 */
};

typedef struct {
  uint32_t status;     /* NUTDB_ST_* */
  uint32_t tok_begin;  /* first significant token of the statement in the token arrays */
  uint32_t tok_count;  /* significant tokens lexed for the statement, incl. the final EOF */
  uint32_t node_begin; /* first node; root = node_begin + node_count - 1 */
  uint32_t node_count; /* 0 unless status == NUTDB_ST_OK */
  uint32_t tok_used;   /* tokens the parser pulled (<= tok_count): what the reference would have lexed */
} NutdbStmt;

typedef struct {
  uint32_t stmt;   /* statement index in the batch */
  uint16_t cls;    /* NUTDB_ST_LEX_ERROR | NUTDB_ST_SYNTAX_ERROR | NUTDB_ST_LIMIT */
  uint16_t code;   /* NUTDB_LE_* for lex errors, NUTDB_SE_* for syntax errors */
  uint32_t line;   /* 1-based; 0 when the variant carries no position */
  uint32_t col;    /* 1-based, in chars; tab = +4 (reference: tokenizer/utf8_iter.rs:89-116) */
  uint32_t pos;    /* byte offset the position was computed from */
  uint32_t a, b, c;
} NutdbError;

/* Output of one batch.  All pointers are HOST pointers (pinned memory) owned by the CONTEXT: they
 * stay valid until nutdb_gpu_batch_free() or the next nutdb_gpu_parse_batch() on the same context,
 * whichever comes first (one live batch per context; the buffers are re-used, grow-only).  The
 * device copies stay alive for the same time (see nutdb_gpu_batch_device). */
typedef struct {
  uint64_t n_stmt;
  uint64_t n_tok;
  uint64_t n_node;
  uint64_t n_err;
  const NutdbStmt *stmt;    /* [n_stmt] */
  const uint8_t *tok_type;  /* [n_tok]  NUTDB_TT_* (no Whitespace / Comment) */
  const uint32_t *tok_start;/* [n_tok]  payload span start, statement-relative */
  const uint32_t *tok_end;  /* [n_tok] */
  const uint8_t *tok_kw;    /* [n_tok]  KeywordOrIdentifier: keyword id or 0; Integer/HexLiteral: digit count (max 255);
                             *          Escaped{SQ,DQ}StringLiteral: 1 = holds no backslash-u escape (unescaping cannot fail); else 0 */
  const NutdbNode *node;    /* [n_node] expanded nodes: NULL in batches produced by the library (see pnode) */
  const NutdbError *err;    /* [n_err]  sorted by .stmt */
  void *impl;               /* opaque */
  const uint32_t *pnode;    /* [n_node] wire nodes (NUTDB_PN_*) */
  uint64_t n_ext;           /* nodes whose fields did not fit the wire word */
  const NutdbNodeExt *ext;  /* [n_ext] sorted by .index; host memory (NULL with NUTDB_F_NO_HOST_COPY until
                               nutdb_gpu_batch_fetch_ext) */
  const uint64_t *wstmt;    /* [n_stmt] with NUTDB_F_WIRE_STMT (then `stmt` is NULL): status | node_count << 4 |
                               tok_used << 34; node_begin is the running sum of node_count (nodes are dense, in statement
                               order).  nutdb_batch_expand_stmts() rebuilds NutdbStmt records from them */
} NutdbBatch;

typedef struct NutdbCtx NutdbCtx;

/* flags for nutdb_gpu_parse_batch */
#define NUTDB_F_NO_TOKENS 1u /* the caller does not want the token arrays (the reference never exposes tokens): no host copy, and the
                                device copies stay in the lexer's segmented layout (statement i's tokens are still
                                [tok_begin, tok_begin + tok_count) of them, but the arrays have gaps between the lexer's ranges) */
#define NUTDB_F_DEVICE_INPUT 2u /* `sql` and `stmt_off` are device pointers on the ctx's device.  No padding is needed behind the last
                                   statement: the kernels read whole 16-byte pieces only below sql + stmt_off[n] and the ragged
                                   rest bytewise.  A text pointer that is not 16-byte aligned costs one device-to-device copy */
#define NUTDB_F_NO_HOST_COPY 4u /* leave every output on the device (use nutdb_gpu_batch_device) */
#define NUTDB_F_WIRE_STMT 16u /* with NUTDB_F_NO_TOKENS: statement records cross PCIe in their 8-byte wire form (NutdbBatch.wstmt)
                                 instead of the 24-byte NutdbStmt -- the token fields are dropped, node_begin is implied */
#define NUTDB_F_OFFSETS32 32u /* `stmt_off` points at n_stmt + 1 uint32_t offsets (cast the pointer): half the upload of the
                                 64-bit form for batches whose text ends below 4 GiB */
#define NUTDB_F_ALL_TOKENS 8u /* lexer verify mode: token arrays also hold Whitespace / Comment tokens (the full stream of
                                 Tokenizer::next_token, tokenizer/mod.rs:66); statements are not parsed */

enum {
  NUTDB_OK = 0,
  NUTDB_E_CUDA = -1,     /* no device / CUDA runtime error (see nutdb_gpu_last_error) */
  NUTDB_E_ARG = -2,      /* bad argument (offsets not ascending, batch too large, ...) */
  NUTDB_E_NOMEM = -3
};

/* Create a context bound to one CUDA device (one process per GPU, or several contexts under the dispatcher
 * below: nutdb_gpu_mctx_*).  Returns NULL on failure. */
NutdbCtx *nutdb_gpu_ctx_create(int device);
void nutdb_gpu_ctx_destroy(NutdbCtx *ctx);
const char *nutdb_gpu_last_error(const NutdbCtx *ctx);

/* Parse a batch: replaces n calls of Parser::parse (src/parser/mod.rs:27).
 * stmt_off has n_stmt+1 ascending entries; total bytes must be < 2^31. */
int nutdb_gpu_parse_batch(NutdbCtx *ctx, const uint8_t *sql, const uint64_t *stmt_off,
                          uint64_t n_stmt, uint32_t flags, NutdbBatch *out);
void nutdb_gpu_batch_free(NutdbCtx *ctx, NutdbBatch *batch);

/* Expands the wire nodes of a batch into NutdbNode records (byte spans, child counts, parent links): pure host
 * arithmetic, no parsing.  `out` must hold n_node records.  Returns 0, or NUTDB_E_ARG. */
int nutdb_batch_expand_nodes(const NutdbBatch *batch, NutdbNode *out);
/* NutdbStmt records of a batch that carries wire statement records (tok_begin / tok_count are 0).  `out`: n_stmt records. */
int nutdb_batch_expand_stmts(const NutdbBatch *batch, NutdbStmt *out);

/* Device-side views of the last batch (valid until batch_free); `node` points at the WIRE nodes (32-bit words). */
typedef struct {
  const void *stmt, *tok_type, *tok_start, *tok_end, *tok_kw, *node, *err;
  const void *wstmt; /* wire statement records (NUTDB_F_WIRE_STMT), else NULL */
} NutdbBatchDevice;
int nutdb_gpu_batch_device(const NutdbBatch *batch, NutdbBatchDevice *out);
/* With NUTDB_F_NO_HOST_COPY a batch reports n_ext but leaves the side table of its wire nodes on the device (ext = NULL);
 * this fetches and sorts it into host memory and sets batch->ext.  A no-op for batches that have it or need none. */
int nutdb_gpu_batch_fetch_ext(NutdbBatch *batch);

/* 64-bit checksum of everything the batch holds on the device (statement records, the four token arrays, nodes,
 * error records), computed on the device:
 *     H = sum over arrays A, 32-bit words (or bytes, for the two byte arrays) w at index i of
 *         mix(mix(i + (A + 1) * 0x9E3779B97F4A7C15) ^ w)    (mod 2^64; mix = the splitmix64 finaliser)
 * with A = 0 stmt, 1 tok_type, 2 tok_start, 3 tok_end, 4 tok_kw, 5 node, 6 err.  Used to compare the outputs of
 * the same log parsed on different numbers of GPUs without moving them to the host. */
int nutdb_gpu_batch_hash(const NutdbBatch *batch, uint64_t *out);

/* Statement splitter for raw buffers (query logs).  The reference has no multi-statement entry
 * (Parser::parse stops after one statement, src/parser/mod.rs:165-172); this produces the `stmt_off`
 * that nutdb_gpu_parse_batch consumes.  The buffer is read as ONE character stream with the
 * tokenizer's rules for '..' / ".." literals (doubled quotes, backslash escapes), `..` identifiers,
 * `--` and block comments (tokenizer/mod.rs:115-184, 313-345, 430-468); every ';' outside of them ends a
 * statement.  Statement i is [stmt_off[i], stmt_off[i+1]) and includes its ';'; white space between
 * statements leads the next one; text after the last ';' is a final statement unless it is only white
 * space.  *stmt_off (n_stmt + 1 entries, stmt_off[0] = 0) is pinned host memory owned by the context,
 * valid until the next call on it.  len < 2^31.  NUTDB_F_DEVICE_INPUT: `sql` is a device pointer. */
int nutdb_gpu_split_statements(NutdbCtx *ctx, const uint8_t *sql, uint64_t len, uint32_t flags,
                               const uint64_t **stmt_off, uint64_t *n_stmt);

/* Single-statement convenience = Parser::parse(sql) (batch of one). */
int nutdb_gpu_parse(NutdbCtx *ctx, const uint8_t *sql, uint64_t len, NutdbBatch *out);

/* Timing of the last parse_batch call, milliseconds measured with CUDA events on the
 * library's stream: [0]=H2D, [1]=lex kernels, [2]=parse kernels, [3]=D2H, [4]=total device. */
int nutdb_gpu_last_timing(const NutdbCtx *ctx, float ms[5]);
/* Number of kernel launches issued by the last parse_batch call. */
int nutdb_gpu_last_launches(const NutdbCtx *ctx);

/* Per-kernel timing: with profiling on, every kernel launch of parse_batch is bracketed by CUDA
 * events on the library's stream.  nutdb_gpu_kernel_timing returns the number of launches of the
 * last call and, for 0 <= i < that number, the kernel's name and duration in milliseconds. */
void nutdb_gpu_set_profiling(NutdbCtx *ctx, int on);
int nutdb_gpu_kernel_timing(const NutdbCtx *ctx, int i, const char **name, float *ms);
/* Statements of the last batch that both table-driven passes declined and the exact automaton
 * parsed (everything malformed, plus constructs outside the tables' grammar). */
uint64_t nutdb_gpu_last_slow_statements(const NutdbCtx *ctx);
/* Statements of the last batch parsed by the second (wide) table-driven pass: deep nesting, arrays, maps, index
 * access, IF, subqueries. */
uint64_t nutdb_gpu_last_wide_statements(const NutdbCtx *ctx);
/* Statements of the last batch that the warp-cooperative lexer handed to the exact walker (every
 * statement with a lex error, hex literals, `$n`, `@name`, code tokens longer than 32 bytes ...). */
uint64_t nutdb_gpu_last_exact_lexed_statements(const NutdbCtx *ctx);
/* The cudaStream_t every kernel and copy of this context is issued on (for callers that want to
 * order their own work or record their own events against it). */
void *nutdb_gpu_ctx_stream(const NutdbCtx *ctx);

/* ----------------------------------------------------------------------------------------
 * Multi-GPU batch dispatcher (north_star: "a multi-GPU batch dispatcher"; statements are independent --
 * Parser::parse keeps no state across calls, src/parser/mod.rs:21-37 -- so a batch shards by statement ranges).
 * One dispatcher drives `n_devices` GPUs of one box from one process: `workers_per_device` contexts per device, one
 * host thread each.  A call parses SHARDS (statement ranges) on their devices and GATHERS each shard's outputs --
 * statement records, wire nodes, error records, and the token arrays unless NUTDB_F_NO_TOKENS -- to one place: pinned
 * host memory (each device copies over its own PCIe link) or, with NUTDB_MF_GATHER_DEVICE0, the memory of the first
 * device (cudaMemcpyPeerAsync: NVLink).  No collective is involved: nothing is reduced.  The consumer gets one
 * callback per shard, from the worker thread that finished it (concurrently from several threads unless
 * NUTDB_MF_SERIAL_CALLBACKS); the views are valid until the callback returns.  Indices inside a chunk (tok_begin,
 * node_begin, err.stmt) are chunk-local; `first_stmt` is the chunk's first statement in the caller's numbering.
 * ---------------------------------------------------------------------------------------- */
typedef struct NutdbMCtx NutdbMCtx;

typedef struct {
  int device_index;        /* position in the dispatcher's device list */
  const uint8_t *sql;      /* text; a device pointer on that device with NUTDB_F_DEVICE_INPUT in `flags` */
  const uint64_t *stmt_off;/* n_stmt + 1 ascending offsets into sql (same memory space as sql) */
  uint64_t n_stmt;
  uint32_t flags;          /* 0 or NUTDB_F_DEVICE_INPUT, NUTDB_F_OFFSETS32 */
  uint64_t first_stmt;     /* handed through to the chunk */
} NutdbMShard;

typedef struct {
  uint64_t shard;          /* index into the shard list */
  uint64_t first_stmt;
  int device;              /* CUDA ordinal that parsed it */
  int on_device;           /* 0: batch pointers are pinned host memory; 1: memory of the dispatcher's first device */
  NutdbBatch batch;        /* counts, stmt, pnode, err (+ tokens); ext is always host memory */
} NutdbMChunk;

typedef void (*nutdb_chunk_fn)(void *user, const NutdbMChunk *chunk);

#define NUTDB_MF_GATHER_DEVICE0 0x100u   /* gather to the first device's memory instead of pinned host memory */
#define NUTDB_MF_SERIAL_CALLBACKS 0x200u /* never run two callbacks at the same time */

NutdbMCtx *nutdb_gpu_mctx_create(const int *devices, int n_devices, int workers_per_device);
void nutdb_gpu_mctx_destroy(NutdbMCtx *m);
const char *nutdb_gpu_mctx_last_error(const NutdbMCtx *m);
int nutdb_gpu_mctx_device_count(const NutdbMCtx *m);
/* Parses the shards on their devices (each device's shards in list order) and gathers.  flags: NUTDB_F_NO_TOKENS,
 * NUTDB_MF_*.  Returns NUTDB_OK or the first error (nutdb_gpu_mctx_last_error). */
int nutdb_gpu_mctx_parse_shards(NutdbMCtx *m, const NutdbMShard *shards, uint64_t n_shards, uint32_t flags,
                                nutdb_chunk_fn fn, void *user);
/* One HOST batch (the arguments of nutdb_gpu_parse_batch; any size; NUTDB_F_OFFSETS32 in `flags` for 32-bit offsets): cut at
 * statement boundaries into chunks of about `chunk_bytes`, contiguous ranges of chunks dealt to the devices balanced by bytes, parsed and gathered as above. */
int nutdb_gpu_mctx_parse_stream(NutdbMCtx *m, const uint8_t *sql, const uint64_t *stmt_off, uint64_t n_stmt,
                                uint64_t chunk_bytes, uint32_t flags, nutdb_chunk_fn fn, void *user);

/* Plain device-to-host copy through the library's CUDA runtime (for hosts without one of their own: reading a chunk
 * that was gathered into device memory, or a nutdb_gpu_batch_device view). */
int nutdb_gpu_copy_to_host(void *dst, const void *src_device, uint64_t bytes);

/* Host-side helpers (pure CPU formatting of results; no parsing):
 * Rust `{:?}` text of statement i's AST, and `Display` text of its error, written
 * NUL-terminated into buf (returns needed length). `sql` must be the statement's own text. */
size_t nutdb_fmt_debug(const NutdbBatch *batch, uint64_t i, const uint8_t *sql, size_t len,
                       char *buf, size_t cap);
size_t nutdb_fmt_error(const NutdbBatch *batch, uint64_t i, const uint8_t *sql, size_t len,
                       char *buf, size_t cap);

const char *nutdb_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NUTDB_GPU_H */
