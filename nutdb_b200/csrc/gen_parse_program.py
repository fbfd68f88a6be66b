#!/usr/bin/env python3
"""Assembler for the parser bytecode (parse_program.h).

The grammar below restates /root/reference/src/parser/mod.rs rule by rule (labels carry the name of
the Rust function they follow, line numbers cited inline).  Calling convention of the pushdown
automaton: `PUSHI arg.. ; CALL f` -> f sees [.. args, ret] and returns with RETN <#args>; CALL pushes
the return pc.  MARK pushes the current node count; CLOSE pops it and emits the interior node over
everything emitted since.  MARKF additionally pushes a slot collecting aux bits (SETAUX) and a `sub`
override (SETSUBX) that CLOSEX consumes.

Run:  python gen_parse_program.py   (rewrites parse_program.h next to this file)
"""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- keyword ids: 1 + index in keyword.rs order (shared with lex_tables.hpp) ----
_src = open(os.path.join(HERE, "lex_tables.hpp")).read()
KEYWORDS = re.findall(r'"([a-z0-9]+)"', _src.split("KEYWORD_TEXT[NUTDB_KW_COUNT] = {")[1].split("};")[0])
assert len(KEYWORDS) == 115
KW = {w.upper(): i + 1 for i, w in enumerate(KEYWORDS)}

# ---- token types / node kinds / lists from include/nutdb_gpu.h ----
_hdr = open(os.path.join(HERE, "..", "..", "include", "nutdb_gpu.h")).read()
CONST = {m.group(1): int(m.group(2)) for m in re.finditer(r"NUTDB_(\w+) = (\d+)", _hdr)}


def C(name):
    return CONST[name]


TT = lambda n: C("TT_" + n)
NK = lambda n: C("NK_" + n)
EL = lambda n: C("EL_" + n)
KL = lambda n: C("KL_" + n)
PF = lambda n: C("PF_" + n)
CF = lambda n: C("CF_" + n)

KWLISTS = {
    "WITH_SELECT": ["with", "select"],
    "ALL_DISTINCT": ["all", "distinct"],
    "ON_USING": ["on", "using"],
    "INSERT_SOURCE": ["values", "from", "select", "with"],
    "TABLE_VIEW": ["table", "view"],
    "TABLE_ATTRS": ["primary", "order", "partition", "comment"],
    "VIEW_ATTRS": ["as", "update", "primary", "order", "partition", "comment"],
    "COLUMN_ATTRS": ["default", "comment"],
    "ALTER_ACTION": ["add", "drop", "rename"],
    "ADD_ENTITY": ["column", "index", "constraint"],
    "DROP_ENTITY": ["column", "index", "constraint", "partition"],
    "RENAME_ENTITY": ["column", "index", "constraint", "table"],
    "DESCRIBE_ENTITY": ["table", "view", "database"],
    "NOT_INFIX": ["in", "like", "ilike", "between", "exists"],
    "NOT_NULL": ["not", "null"],
    "INTERVAL_UNIT": ["second", "minute", "hour", "day", "month", "year"],
    "DATATYPE": ["int8", "int16", "int32", "int64", "int128", "uint8", "uint16", "uint32", "uint64", "uint128",
                 "serial32", "serial64", "serial128", "userial32", "userial64", "userial128", "decimal32", "decimal64",
                 "float32", "float64", "boolean", "chars", "string", "uuid", "date", "datetime", "array", "enum",
                 "tuple", "map", "dictionary", "nullable"],
    "CASE_NEXT": ["when", "else", "end"],
    "VIEW_NEEDS_UPDATE": ["update"],
}
EXPECTED = {
    "RParen": ["RParen"], "LParen": ["LParen"], "RBracket": ["RBracket"], "RBrace": ["RBrace"],
    "NegLiteral": ["IntegerLiteral", "HexLiteral", "FloatLiteral"],
    "Identifier": ["DelimitedIdentifier", "KeywordOrIdentifier", "Mul"],
    "Colon": ["Colon"], "Keyword": ["KeywordOrIdentifier"],
    "IdentString": ["KeywordOrIdentifier", "DelimitedIdentifier"],
    "IntLiteral": ["IntegerLiteral", "HexLiteral"],
    "StrLiteral": ["RawStringLiteral", "EscapedSQStringLiteral", "EscapedDQStringLiteral"],
    "ConfigIdent": ["ConfigIdentifier"], "Eq": ["Eq"], "Comma": ["Comma"],
    "Prefix": ["RawStringLiteral", "EscapedSQStringLiteral", "EscapedDQStringLiteral", "FloatLiteral", "HexLiteral",
               "IntegerLiteral", "QueryParameter", "KeywordOrIdentifier", "DelimitedIdentifier", "LParen", "LBracket",
               "LBrace", "Minus", "Plus", "BitNot", "Mul"],
}

# op name -> number of extra words ('T' = jump-table whose length the op knows)
OPS = [
    ("EXPECT", 0), ("IFTT", 1), ("PEEKTT", 1), ("PEEKTERM", 1), ("KW", 0), ("TRYKW", 1), ("PEEKKW", 1), ("PEEKKW2", 2),
    ("ADV", 0), ("ONEOF", "T"), ("CALL", 1), ("RET", 0), ("RETN", 0), ("JMP", 1), ("PUSHI", 0), ("POP", 0), ("PICK", 0),
    ("MARK", 0), ("MARKF", 0), ("SETAUX", 0), ("SETSUBX", 0), ("CLOSE", 1), ("CLOSEX", 0), ("LEAFTOK", 1),
    ("LEAFTOKB", 1), ("LEAF0", 1), ("SETSUB", 0), ("INTLIT", 1), ("STRLIT", 0), ("SAVETOK", 0), ("CHECK_SUBQUERY", 0),
    ("CHECK_SOURCE", 0), ("CHECK_FNCALL", 0), ("DUPCHECK", 1), ("NEED_STRATEGY", 0), ("PUSH_LASTB", 0),
    ("CHECK_WIDTH", 0), ("TUPLE_OR_SINGLE", 0), ("PREFIX_DISPATCH", "T"), ("POWER", 1), ("INFIX_DISPATCH", "T"),
    ("KWINFIX", "T"), ("PUSHR0", 0), ("NODE_BIN", 0), ("NODE", 1), ("FOLD", 0), ("IDENT_PREFIX", 0), ("IDENTIFIER", 0),
    ("FAILAT", 1), ("UPOWER", 1), ("UTYPE", 0), ("NODE_UNION", 0), ("JOINHEAD", 1), ("STMT_DISPATCH", "T"),
    ("FINISH", 0), ("DROPLEFT", 0),
]
OPNUM = {n: i + 1 for i, (n, _) in enumerate(OPS)}
OPEXTRA = dict(OPS)

code = []     # list of ints or label-name strings
labels = {}


def L(name):
    assert name not in labels, name
    labels[name] = len(code)


def op(name, arg=0, *words):
    assert 0 <= arg < 256, (name, arg)
    extra = OPEXTRA[name]
    if extra != "T":
        assert len(words) == extra, (name, words)
    code.append(OPNUM[name] | (arg << 8))
    code.extend(words)


# binary operator ordinals (ast/item.rs:136-164) and friends
B = dict(Plus=0, Minus=1, Multi=2, Div=3, Mod=4, Gt=5, Lt=6, GtEq=7, LtEq=8, Eq=9, NotEq=10, And=11, Or=12, Xor=13,
         Like=14, NotLike=15, ILike=16, NotILike=17, In=18, NotIn=19, IndexAccess=20)
F = dict(If=0, MultiIf=1, CaseWhen=2, Between=3, NotBetween=4, Exists=5, NotExists=6, Others=7)
P = dict(Terminator=0, Or=1, Xor=2, And=3, Not=4, Comparison=5, Between=6)
WIDTH_U8, WIDTH_USIZE, WIDTH_U128 = 0, 1, 2


def ident_string(kind, sub=0):  # must_parse_identifier_string (mod.rs:1682) -> leaf
    op("EXPECT", EL("IdentString"))
    op("LEAFTOK", NK(kind), sub)


def if_exists(kws, after):  # `if self.try_parse_keyword(IF)? { must_parse_keyword(s)...; true }`
    op("TRYKW", KW["IF"], after)
    for k in kws:
        op("KW", KW[k])
    op("SETAUX", 1)
    L(after)


# =============================================================================================
# parse_stmt (mod.rs:128-180)
# =============================================================================================
L("STMT")
op("STMT_DISPATCH", 0, "S_WITH", "S_SELECT", "S_INSERT", "S_EXPLAIN", "S_ALTER", "S_CREATE", "S_DESCRIBE", "S_DROP",
   "S_TRUNCATE", "S_OPTIMIZE", "S_SET", "STMT_END")
L("STMT_END")
op("FINISH")

# try_parse_select_stmt (mod.rs:190-203)
L("S_SELECT")
op("MARK"); op("PUSHI", 0); op("CALL", 0, "QUERY_TDOP_SELECT"); op("CLOSE", NK("STMT_SELECT"), 0); op("JMP", 0, "STMT_END")
L("S_WITH")
op("MARK"); op("PUSHI", 0); op("CALL", 0, "QUERY_TDOP_WITH"); op("CLOSE", NK("STMT_SELECT"), 0); op("JMP", 0, "STMT_END")

# must_parse_subquery (mod.rs:206) / must_parse_subquery_tdop (mod.rs:218-241): [power, ret]
L("SUBQUERY")
op("PUSHI", 0); op("CALL", 0, "SUBQ_TDOP"); op("RET")
L("SUBQ_TDOP")
op("IFTT", TT("LParen"), "SQ_NOPAREN")
op("ONEOF", KL("WITH_SELECT"), "SQP_WITH", "SQP_SELECT")
L("SQP_WITH")
op("PUSHI", 0); op("CALL", 0, "QUERY_TDOP_WITH"); op("EXPECT", EL("RParen")); op("RETN", 1)
L("SQP_SELECT")
op("PUSHI", 0); op("CALL", 0, "QUERY_TDOP_SELECT"); op("EXPECT", EL("RParen")); op("RETN", 1)
L("SQ_NOPAREN")
op("ONEOF", KL("WITH_SELECT"), "SQN_WITH", "SQN_SELECT")
L("SQN_WITH")
op("PICK", 1); op("CALL", 0, "QUERY_TDOP_WITH"); op("RETN", 1)
L("SQN_SELECT")
op("PICK", 1); op("CALL", 0, "QUERY_TDOP_SELECT"); op("RETN", 1)

# must_parse_query_tdop (mod.rs:243-276): [power, ret]
L("QUERY_TDOP_WITH")
op("CALL", 0, "BODY_WITH"); op("JMP", 0, "QT_LOOP")
L("QUERY_TDOP_SELECT")
op("CALL", 0, "BODY_SELECT")
L("QT_LOOP")
op("UPOWER", 0, "QT_END")       # -> [power, ret, np]
op("UTYPE")                     # -> [power, ret, np, typ]
op("PICK", 1)                   # np
op("CALL", 0, "SUBQ_TDOP")
op("NODE_UNION")                # pops typ, np
op("JMP", 0, "QT_LOOP")
L("QT_END")
op("RETN", 1)

# must_parse_query_body (mod.rs:279-325)
L("BODY_WITH")
op("MARK")
op("MARK")                       # must_parse_query_clause_with (mod.rs:327-347)
L("BW_LOOP")
ident_string("NAME")
op("KW", KW["AS"])
op("SAVETOK")
op("CALL", 0, "EXPR")
op("CHECK_SUBQUERY")
op("IFTT", TT("Comma"), "BW_END"); op("JMP", 0, "BW_LOOP")
L("BW_END")
op("CLOSE", NK("WITH"), 0)
op("KW", KW["SELECT"])
op("JMP", 0, "BODY_COMMON")
L("BODY_SELECT")
op("MARK")
L("BODY_COMMON")
op("TRYKW", KW["DISTINCT"], "B_COLS")
op("MARKF")                      # must_parse_query_clause_distinct (mod.rs:349-360)
op("TRYKW", KW["ON"], "B_DIST_CLOSE")
op("SETAUX", 1)
op("EXPECT", EL("LParen")); op("CALL", 0, "QEXPR_LIST"); op("EXPECT", EL("RParen"))
L("B_DIST_CLOSE")
op("CLOSEX", NK("DISTINCT"))
L("B_COLS")
op("MARK"); op("CALL", 0, "QEXPR_LIST"); op("CLOSE", NK("COLS"), 0)
op("PEEKKW", KW["FROM"], "B_JOINS")          # try_parse_query_clause_from (mod.rs:362-374)
op("ADV"); op("MARK"); op("CALL", 0, "QSOURCE"); op("CLOSE", NK("FROM"), 0)
L("B_JOINS")
op("JOINHEAD", 0, "B_WHERE")                  # try_parse_query_clause_join (mod.rs:376-431)
op("KW", KW["JOIN"])
op("CALL", 0, "QSOURCE")
op("ONEOF", KL("ON_USING"), "J_ON", "J_USING")
L("J_ON")
op("CALL", 0, "EXPR"); op("CLOSEX", NK("JOIN")); op("JMP", 0, "B_JOINS")
L("J_USING")
op("SETAUX", 1); op("EXPECT", EL("LParen"))
L("JU_LOOP")
op("IDENTIFIER"); op("IFTT", TT("Comma"), "JU_END"); op("JMP", 0, "JU_LOOP")
L("JU_END")
op("EXPECT", EL("RParen")); op("CLOSEX", NK("JOIN")); op("JMP", 0, "B_JOINS")
L("B_WHERE")                                   # mod.rs:433-445
op("PEEKKW", KW["WHERE"], "B_GROUP"); op("ADV"); op("MARK"); op("CALL", 0, "EXPR"); op("CLOSE", NK("WHERE"), 0)
L("B_GROUP")                                   # mod.rs:447-460
op("PEEKKW", KW["GROUP"], "B_HAVING"); op("ADV"); op("KW", KW["BY"]); op("MARK"); op("CALL", 0, "QEXPR_LIST")
op("CLOSE", NK("GROUPBY"), 0)
L("B_HAVING")                                  # mod.rs:462-474
op("PEEKKW", KW["HAVING"], "B_ORDER"); op("ADV"); op("MARK"); op("CALL", 0, "EXPR"); op("CLOSE", NK("HAVING"), 0)
L("B_ORDER")                                   # mod.rs:476-501 (DESC is tested twice, ASC never)
op("PEEKKW", KW["ORDER"], "B_LIMIT"); op("ADV"); op("KW", KW["BY"]); op("MARK")
L("BO_LOOP")
op("CALL", 0, "QEXPR")
op("TRYKW", KW["DESC"], "BO_ASC"); op("LEAF0", NK("ORDER_DESC"), 0); op("JMP", 0, "BO_NEXT")
L("BO_ASC")
op("TRYKW", KW["DESC"], "BO_NEXT")
L("BO_NEXT")
op("IFTT", TT("Comma"), "BO_END"); op("JMP", 0, "BO_LOOP")
L("BO_END")
op("CLOSE", NK("ORDERBY"), 0)
L("B_LIMIT")                                   # mod.rs:503-544
op("PEEKKW", KW["LIMIT"], "B_END"); op("ADV"); op("MARKF")
op("INTLIT", NK("NUM"), WIDTH_USIZE)
op("IFTT", TT("Comma"), "BL_KW")
op("SETSUBX", 1); op("INTLIT", NK("NUM"), WIDTH_USIZE); op("JMP", 0, "BL_TIES")
L("BL_KW")
op("TRYKW", KW["OFFSET"], "BL_TIES")
op("SETSUBX", 2); op("INTLIT", NK("NUM"), WIDTH_USIZE)
L("BL_TIES")
op("TRYKW", KW["WITH"], "BL_CLOSE"); op("KW", KW["TIES"]); op("SETAUX", 1)
L("BL_CLOSE")
op("CLOSEX", NK("LIMIT"))
L("B_END")
op("CLOSE", NK("QUERY_BODY"), 0)
op("RET")

# must_parse_query_source (mod.rs:546-569)
L("QSOURCE")
op("SAVETOK"); op("CALL", 0, "EXPR"); op("CHECK_SOURCE")
op("TRYKW", KW["AS"], "QS_END"); ident_string("ALIAS")
L("QS_END")
op("RET")
# must_parse_query_expr (mod.rs:571-579) / _list (:581-585)
L("QEXPR")
op("CALL", 0, "EXPR"); op("TRYKW", KW["AS"], "QE_END"); ident_string("ALIAS")
L("QE_END")
op("RET")
L("QEXPR_LIST")
op("CALL", 0, "QEXPR"); op("IFTT", TT("Comma"), "QEL_END"); op("JMP", 0, "QEXPR_LIST")
L("QEL_END")
op("RET")

# =============================================================================================
# INSERT (mod.rs:589-670)
# =============================================================================================
L("S_INSERT")
op("MARK"); op("KW", KW["INTO"]); ident_string("NAME")
op("IFTT", TT("LParen"), "SI_SRC")
L("SI_COLS")
ident_string("NAME"); op("IFTT", TT("Comma"), "SI_COLS_END"); op("JMP", 0, "SI_COLS")
L("SI_COLS_END")
op("EXPECT", EL("RParen"))
L("SI_SRC")
op("SAVETOK")
op("ONEOF", KL("INSERT_SOURCE"), "SI_VALUES", "SI_FROM", "SI_SELECT", "SI_WITH")
L("SI_VALUES")
op("POP"); op("CALL", 0, "ROWS"); op("JMP", 0, "SI_END")
L("SI_FROM")
op("CALL", 0, "EXPR"); op("CHECK_FNCALL", PF("INSERT_SOURCE")); op("JMP", 0, "SI_END")
L("SI_SELECT")
op("POP"); op("PUSHI", 0); op("CALL", 0, "QUERY_TDOP_SELECT"); op("JMP", 0, "SI_END")
L("SI_WITH")
op("POP"); op("PUSHI", 0); op("CALL", 0, "QUERY_TDOP_WITH")
L("SI_END")
op("CLOSE", NK("STMT_INSERT"), 0); op("JMP", 0, "STMT_END")
# must_parse_insert_rows (mod.rs:636-670)
L("ROWS")
op("MARK"); op("EXPECT", EL("LParen")); op("MARK")
L("R1")
op("CALL", 0, "EXPR"); op("IFTT", TT("Comma"), "R1E"); op("JMP", 0, "R1")
L("R1E")
op("CLOSE", NK("ROW"), 0); op("EXPECT", EL("RParen")); op("PUSH_LASTB")
op("IFTT", TT("Comma"), "ROWS_END")
L("RL")
op("EXPECT", EL("LParen")); op("MARK")
L("R2")
op("CALL", 0, "EXPR"); op("IFTT", TT("Comma"), "R2E"); op("JMP", 0, "R2")
L("R2E")
op("CLOSE", NK("ROW"), 0); op("CHECK_WIDTH"); op("EXPECT", EL("RParen"))
op("IFTT", TT("Comma"), "ROWS_END"); op("JMP", 0, "RL")
L("ROWS_END")
op("POP"); op("CLOSE", NK("ROWS"), 0); op("RET")

# EXPLAIN (mod.rs:674-685)
L("S_EXPLAIN")
op("MARK"); op("CALL", 0, "SUBQUERY"); op("CLOSE", NK("STMT_EXPLAIN"), 0); op("JMP", 0, "STMT_END")

# =============================================================================================
# CREATE (mod.rs:689-972)
# =============================================================================================
L("S_CREATE")
op("MARKF")
op("ONEOF", KL("TABLE_VIEW"), "SC_T", "SC_V")
L("SC_T")
if_exists(["NOT", "EXISTS"], "SC_T2")
op("CALL", 0, "TABLEDEF"); op("CLOSEX", NK("STMT_CREATE")); op("JMP", 0, "STMT_END")
L("SC_V")
if_exists(["NOT", "EXISTS"], "SC_V2")
op("CALL", 0, "VIEWDEF"); op("CLOSEX", NK("STMT_CREATE")); op("JMP", 0, "STMT_END")

# must_parse_table_definition (mod.rs:712-805)
L("TABLEDEF")
op("MARK"); ident_string("NAME"); op("EXPECT", EL("LParen"))
L("TD_ITEM")
op("TRYKW", KW["INDEX"], "TD_I2"); op("CALL", 0, "INDEXDEF"); op("JMP", 0, "TD_NEXT")
L("TD_I2")
op("TRYKW", KW["CONSTRAINT"], "TD_I3"); op("CALL", 0, "CONSTRDEF"); op("JMP", 0, "TD_NEXT")
L("TD_I3")
op("CALL", 0, "COLDEF")
L("TD_NEXT")
op("IFTT", TT("Comma"), "TD_ITEMS_END"); op("JMP", 0, "TD_ITEM")
L("TD_ITEMS_END")
op("EXPECT", EL("RParen")); op("PUSHI", 0)
L("TD_ATTR")
op("PEEKTT", TT("KeywordOrIdentifier"), "TD_END")
op("SAVETOK")
op("ONEOF", KL("TABLE_ATTRS"), "TA_PK", "TA_ORDER", "TA_PART", "TA_COMMENT")
L("TA_PK")
op("DUPCHECK", 0, CF("PRIMARY_KEY")); op("KW", KW["KEY"]); op("MARK"); op("CALL", 0, "EXPR_LIST")
op("CLOSE", NK("ATTR_PK"), 0); op("JMP", 0, "TD_ATTR")
L("TA_ORDER")
op("DUPCHECK", 1, CF("ORDER_BY")); op("KW", KW["BY"]); op("MARK"); op("CALL", 0, "EXPR_LIST")
op("CLOSE", NK("ATTR_ORDER"), 0); op("JMP", 0, "TD_ATTR")
L("TA_PART")
op("DUPCHECK", 2, CF("PARTITION_BY")); op("KW", KW["BY"]); op("MARK"); op("CALL", 0, "EXPR")
op("CLOSE", NK("ATTR_PART"), 0); op("JMP", 0, "TD_ATTR")
L("TA_COMMENT")
op("DUPCHECK", 3, CF("COMMENT")); op("STRLIT"); op("JMP", 0, "TD_ATTR")
L("TD_END")
op("POP"); op("CLOSE", NK("TABLEDEF"), 0); op("RET")

# must_parse_view_definition (mod.rs:807-911)
L("VIEWDEF")
op("MARK"); ident_string("NAME"); op("PUSHI", 0)
L("VD_ATTR")
op("SAVETOK")
op("ONEOF", KL("VIEW_ATTRS"), "VA_AS", "VA_UPDATE", "VA_PK", "VA_ORDER", "VA_PART", "VA_COMMENT")
L("VA_AS")
op("NEED_STRATEGY"); op("POP"); op("CALL", 0, "SUBQUERY"); op("CLOSE", NK("VIEWDEF"), 0); op("RET")
L("VA_UPDATE")
op("DUPCHECK", 0, CF("UPDATE_BY")); op("KW", KW["BY"]); ident_string("STRATEGY"); op("JMP", 0, "VD_ATTR")
L("VA_PK")
op("DUPCHECK", 1, CF("PRIMARY_KEY")); op("KW", KW["KEY"]); op("MARK"); op("CALL", 0, "EXPR_LIST")
op("CLOSE", NK("ATTR_PK"), 0); op("JMP", 0, "VD_ATTR")
L("VA_ORDER")
op("DUPCHECK", 2, CF("ORDER_BY")); op("KW", KW["BY"]); op("MARK"); op("CALL", 0, "EXPR_LIST")
op("CLOSE", NK("ATTR_ORDER"), 0); op("JMP", 0, "VD_ATTR")
L("VA_PART")
op("DUPCHECK", 3, CF("PARTITION_BY")); op("KW", KW["BY"]); op("MARK"); op("CALL", 0, "EXPR")
op("CLOSE", NK("ATTR_PART"), 0); op("JMP", 0, "VD_ATTR")
L("VA_COMMENT")
op("DUPCHECK", 4, CF("COMMENT")); op("STRLIT"); op("JMP", 0, "VD_ATTR")

# must_parse_constraint_def (mod.rs:913-918) / must_parse_index_def (:920-934) / must_parse_column_def (:936-972)
L("CONSTRDEF")
op("MARK"); ident_string("NAME"); op("KW", KW["CHECK"]); op("CALL", 0, "EXPR"); op("CLOSE", NK("CONSTRDEF"), 0); op("RET")
L("INDEXDEF")
op("MARK"); ident_string("NAME"); op("SAVETOK"); op("CALL", 0, "EXPR"); op("CHECK_FNCALL", PF("INDEXER"))
op("CLOSE", NK("INDEXDEF"), 0); op("RET")
L("COLDEF")
op("MARK"); ident_string("NAME"); op("CALL", 0, "DATATYPE"); op("PUSHI", 0)
L("CD_ATTR")
op("PEEKTT", TT("KeywordOrIdentifier"), "CD_END")
op("SAVETOK")
op("ONEOF", KL("COLUMN_ATTRS"), "CA_DEFAULT", "CA_COMMENT")
L("CA_DEFAULT")
op("DUPCHECK", 0, CF("DEFAULT")); op("MARK"); op("CALL", 0, "EXPR"); op("CLOSE", NK("ATTR_DEFAULT"), 0)
op("JMP", 0, "CD_ATTR")
L("CA_COMMENT")
op("DUPCHECK", 1, CF("COMMENT")); op("STRLIT"); op("JMP", 0, "CD_ATTR")
L("CD_END")
op("POP"); op("CLOSE", NK("COLDEF"), 0); op("RET")

# must_parse_datatype (mod.rs:1688-1797)
L("DATATYPE")
targets = []
for i in range(32):
    targets.append("DT_%d" % i)
op("ONEOF", KL("DATATYPE"), *targets)
for i in range(26):
    L("DT_%d" % i)
    if i in (16, 17, 21):
        op("MARK"); op("EXPECT", EL("LParen")); op("INTLIT", NK("NUM"), WIDTH_USIZE if i == 21 else WIDTH_U8)
        op("EXPECT", EL("RParen")); op("CLOSE", NK("DT_PARAM"), i); op("RET")
    elif i == 22:
        op("PEEKTT", TT("LParen"), "DT_STR0")
        op("MARK"); op("EXPECT", EL("LParen")); op("INTLIT", NK("NUM"), WIDTH_USIZE); op("EXPECT", EL("RParen"))
        op("CLOSE", NK("DT_PARAM"), 22); op("RET")
        L("DT_STR0")
        op("LEAF0", NK("DT_SCALAR"), 22); op("RET")
    else:
        op("LEAF0", NK("DT_SCALAR"), i); op("RET")
for i, sub in ((26, 0), (30, 4), (31, 5)):   # Array / Dictionary / Nullable
    L("DT_%d" % i)
    op("MARK"); op("EXPECT", EL("LParen")); op("CALL", 0, "DATATYPE"); op("EXPECT", EL("RParen"))
    op("CLOSE", NK("DT_COMPOUND"), sub); op("RET")
L("DT_27")                                       # Enum: must_parse_enum_binds (mod.rs:1799-1813)
op("MARK"); op("EXPECT", EL("LParen"))
L("EN_LOOP")
op("STRLIT"); op("IFTT", TT("Eq"), "EN_NOID"); op("INTLIT", NK("NUM"), WIDTH_USIZE)
L("EN_NOID")
op("IFTT", TT("Comma"), "EN_END"); op("JMP", 0, "EN_LOOP")
L("EN_END")
op("EXPECT", EL("RParen")); op("CLOSE", NK("DT_COMPOUND"), 1); op("RET")
L("DT_28")                                       # Tuple
op("MARK"); op("EXPECT", EL("LParen"))
L("TU_LOOP")
op("CALL", 0, "DATATYPE"); op("IFTT", TT("Comma"), "TU_END"); op("JMP", 0, "TU_LOOP")
L("TU_END")
op("EXPECT", EL("RParen")); op("CLOSE", NK("DT_COMPOUND"), 2); op("RET")
L("DT_29")                                       # Map(key, value) -- children stay in source order
op("MARK"); op("EXPECT", EL("LParen")); op("CALL", 0, "DATATYPE"); op("EXPECT", EL("Comma")); op("CALL", 0, "DATATYPE")
op("EXPECT", EL("RParen")); op("CLOSE", NK("DT_COMPOUND"), 3); op("RET")

# =============================================================================================
# ALTER (mod.rs:976-1059)
# =============================================================================================
L("S_ALTER")
op("MARKF"); op("KW", KW["TABLE"]); ident_string("NAME")
op("ONEOF", KL("ALTER_ACTION"), "AL_ADD", "AL_DROP", "AL_RENAME")
L("AL_ADD")
if_exists(["NOT", "EXISTS"], "AA1")
op("ONEOF", KL("ADD_ENTITY"), "AA_COL", "AA_IDX", "AA_CON")
L("AA_COL")
op("CALL", 0, "COLDEF"); op("JMP", 0, "AA_POS")
L("AA_IDX")
op("CALL", 0, "INDEXDEF"); op("JMP", 0, "AA_POS")
L("AA_CON")
op("CALL", 0, "CONSTRDEF")
L("AA_POS")
op("TRYKW", KW["FIRST"], "AA_P2"); op("LEAF0", NK("POS_FIRST"), 0); op("JMP", 0, "AA_END")
L("AA_P2")
op("TRYKW", KW["AFTER"], "AA_END"); ident_string("POS_AFTER")
L("AA_END")
op("CLOSEX", NK("STMT_ALTER")); op("JMP", 0, "STMT_END")
L("AL_DROP")
op("SETSUBX", 1)
if_exists(["EXISTS"], "AD1")
op("ONEOF", KL("DROP_ENTITY"), "AD_C", "AD_I", "AD_K", "AD_P")
for lab, sub in (("AD_C", 0), ("AD_I", 1), ("AD_K", 2)):
    L(lab)
    ident_string("ENT_NAME", sub); op("JMP", 0, "AD_END")
L("AD_P")
op("STRLIT")
L("AD_END")
op("CLOSEX", NK("STMT_ALTER")); op("JMP", 0, "STMT_END")
L("AL_RENAME")
op("SETSUBX", 2)
op("ONEOF", KL("RENAME_ENTITY"), "AR_C", "AR_I", "AR_K", "AR_T")
for lab, sub in (("AR_C", 0), ("AR_I", 1), ("AR_K", 2)):
    L(lab)
    ident_string("ENT_NAME", sub); op("JMP", 0, "AR_NEW")
L("AR_T")
op("LEAF0", NK("ENT_NAME"), 3)
L("AR_NEW")
ident_string("NAME"); op("CLOSEX", NK("STMT_ALTER")); op("JMP", 0, "STMT_END")

# DESCRIBE (mod.rs:1063-1079)
L("S_DESCRIBE")
op("MARK")
op("ONEOF", KL("DESCRIBE_ENTITY"), "D_T", "D_V", "D_D")
L("D_T")
ident_string("NAME"); op("CLOSE", NK("STMT_DESCRIBE"), 0); op("JMP", 0, "STMT_END")
L("D_V")
ident_string("NAME"); op("CLOSE", NK("STMT_DESCRIBE"), 1); op("JMP", 0, "STMT_END")
L("D_D")
op("CLOSE", NK("STMT_DESCRIBE"), 2); op("JMP", 0, "STMT_END")

# DROP (mod.rs:1083-1110) / TRUNCATE (:1114-1141)
for name, kind in (("S_DROP", "STMT_DROP"), ("S_TRUNCATE", "STMT_TRUNCATE")):
    L(name)
    op("MARKF")
    op("ONEOF", KL("TABLE_VIEW"), name + "_T", name + "_V")
    L(name + "_V")
    op("SETSUBX", 1)
    L(name + "_T")
    if_exists(["EXISTS"], name + "_N")
    ident_string("NAME"); op("CLOSEX", NK(kind)); op("JMP", 0, "STMT_END")

# OPTIMIZE (mod.rs:1146-1171)
L("S_OPTIMIZE")
op("MARK"); op("KW", KW["TABLE"]); ident_string("NAME")
op("PEEKTERM", 0, "SO_END")
op("KW", KW["ON"]); op("KW", KW["PARTITION"]); op("CALL", 0, "EXPR")
L("SO_END")
op("CLOSE", NK("STMT_OPTIMIZE"), 0); op("JMP", 0, "STMT_END")

# SET (mod.rs:1176-1195)
L("S_SET")
op("MARK"); op("EXPECT", EL("ConfigIdent")); op("LEAFTOK", NK("NAME"), 0); op("EXPECT", EL("Eq")); op("CALL", 0, "EXPR")
op("CLOSE", NK("STMT_SET"), 0); op("JMP", 0, "STMT_END")

# =============================================================================================
# expressions (mod.rs:1198-1619)
# =============================================================================================
L("EXPR_LIST")                                   # must_parse_expr_list (mod.rs:1199-1203)
op("CALL", 0, "EXPR"); op("IFTT", TT("Comma"), "EL_END"); op("JMP", 0, "EXPR_LIST")
L("EL_END")
op("RET")
L("EXPR")                                        # must_parse_expr (mod.rs:1205)
op("PUSHI", P["Terminator"]); op("CALL", 0, "TDOP"); op("RET")
L("TDOP")                                        # must_parse_expr_tdop (mod.rs:1209-1220): [power, ret]
op("CALL", 0, "PREFIX")
L("TDOP_LOOP")
op("POWER", 0, "TDOP_END")                       # -> [power, ret, this_power]
op("CALL", 0, "INFIX")                           # INFIX: [.., this_power, ret2], returns RETN 1
op("JMP", 0, "TDOP_LOOP")
L("TDOP_END")
op("RETN", 1)

L("PREFIX")                                      # must_parse_expr_prefix (mod.rs:1222-1347)
op("PREFIX_DISPATCH", 0, "PX_DONE", "PX_LPAREN", "PX_LBRACKET", "PX_LBRACE", "PREFIX", "PX_BITNOT", "PX_NOT",
   "PX_INTERVAL", "PX_IF", "PX_CASE", "PX_WORD", "PX_DELIM", "PX_QPARAM")
L("PX_DONE")
op("RET")
L("PX_BITNOT")
op("CALL", 0, "PREFIX"); op("NODE", NK("UNARY"), 0 | (1 << 8)); op("RET")
L("PX_NOT")
op("CALL", 0, "PREFIX"); op("FOLD", 5); op("RET")
L("PX_LPAREN")                                   # mod.rs:1229-1246
op("PEEKKW2", KW["SELECT"], KW["WITH"], "PX_PAREN_SUBQ")
op("MARK"); op("CALL", 0, "EXPR_LIST"); op("TUPLE_OR_SINGLE"); op("EXPECT", EL("RParen")); op("RET")
L("PX_PAREN_SUBQ")
op("CALL", 0, "SUBQUERY"); op("EXPECT", EL("RParen")); op("RET")
L("PX_LBRACKET")
op("MARK"); op("CALL", 0, "EXPR_LIST"); op("CLOSE", NK("COLLECTION"), 2); op("EXPECT", EL("RBracket")); op("RET")
L("PX_LBRACE")                                   # must_parse_map (mod.rs:1558-1568)
op("MARK")
L("MAP_LOOP")
op("CALL", 0, "EXPR"); op("EXPECT", EL("Colon")); op("CALL", 0, "EXPR"); op("IFTT", TT("Comma"), "MAP_END")
op("JMP", 0, "MAP_LOOP")
L("MAP_END")
op("CLOSE", NK("COLLECTION"), 1); op("EXPECT", EL("RBrace")); op("RET")
L("PX_INTERVAL")                                 # must_parse_interval (mod.rs:1489-1503)
op("INTLIT", NK("LIT_INTERVAL"), WIDTH_USIZE)
op("ONEOF", KL("INTERVAL_UNIT"), "IU0", "IU1", "IU2", "IU3", "IU4", "IU5")
for i in range(6):
    L("IU%d" % i)
    op("SETSUB", i); op("RET")
L("PX_IF")                                       # must_parse_if_body (mod.rs:1571-1582)
op("MARK"); op("CALL", 0, "EXPR"); op("KW", KW["THEN"]); op("CALL", 0, "EXPR"); op("KW", KW["ELSE"])
op("CALL", 0, "EXPR"); op("KW", KW["END"]); op("CLOSE", NK("FNCALL"), F["If"]); op("RET")
L("PX_CASE")                                     # must_parse_case_when_body (mod.rs:1585-1618)
op("MARKF")
op("TRYKW", KW["WHEN"], "CS_SCRUT")
op("SETSUBX", F["MultiIf"]); op("JMP", 0, "CS_BRANCH")
L("CS_SCRUT")
op("CALL", 0, "EXPR"); op("KW", KW["WHEN"]); op("SETSUBX", F["CaseWhen"])
L("CS_BRANCH")
op("CALL", 0, "EXPR"); op("KW", KW["THEN"]); op("CALL", 0, "EXPR")
op("ONEOF", KL("CASE_NEXT"), "CS_BRANCH", "CS_ELSE", "CS_END")
L("CS_ELSE")
op("CALL", 0, "EXPR"); op("KW", KW["END"]); op("JMP", 0, "CS_CLOSE")
L("CS_END")
op("LEAF0", NK("LIT_NULL"), 0)
L("CS_CLOSE")
op("CLOSEX", NK("FNCALL")); op("RET")
L("PX_WORD")                                     # mod.rs:1303-1308
op("IFTT", TT("LParen"), "PX_IDENT")
op("MARK"); op("LEAFTOKB", NK("FN_NAME"), 2); op("CALL", 0, "FNARGS_REST"); op("CLOSE", NK("FNCALL"), F["Others"])
op("RET")
L("PX_IDENT")
op("IDENT_PREFIX"); op("RET")
L("PX_DELIM")
op("IDENT_PREFIX"); op("RET")
L("PX_QPARAM")                                   # mod.rs:1311 (sic: the index comes from a SECOND integer token)
op("INTLIT", NK("QPARAM"), WIDTH_USIZE); op("RET")
L("FNARGS_REST")                                 # try_parse_fn_call_args after '(' (mod.rs:1538-1556)
op("PEEKTT", TT("RParen"), "FA1"); op("ADV"); op("RET")
L("FA1")
op("PEEKKW2", KW["SELECT"], KW["WITH"], "FA_SUBQ")
op("CALL", 0, "EXPR_LIST"); op("EXPECT", EL("RParen")); op("RET")
L("FA_SUBQ")
op("CALL", 0, "SUBQUERY"); op("EXPECT", EL("RParen")); op("RET")

L("INFIX")                                       # must_parse_expr_infix (mod.rs:1349-1486): [this_power, ret]
op("INFIX_DISPATCH", 0, "IX_BIN", "IX_EQ", "IX_NEQ", "IX_INDEX", "IX_KW")
L("IX_BIN")
op("PUSHR0"); op("PICK", 2); op("CALL", 0, "TDOP"); op("NODE_BIN"); op("RETN", 1)
L("IX_EQ")
op("PICK", 1); op("CALL", 0, "TDOP"); op("FOLD", 0); op("RETN", 1)
L("IX_NEQ")
op("PICK", 1); op("CALL", 0, "TDOP"); op("FOLD", 1); op("RETN", 1)
L("IX_INDEX")
op("CALL", 0, "EXPR"); op("EXPECT", EL("RBracket")); op("NODE", NK("BINARY"), B["IndexAccess"] | (2 << 8))
op("RETN", 1)
L("IX_KW")
op("KWINFIX", 0, "IX_AND", "IX_OR", "IX_XOR", "IX_NOT", "IX_IS", "IX_BIN", "IX_BETWEEN")
L("IX_AND")
op("PICK", 1); op("CALL", 0, "TDOP"); op("FOLD", 2); op("RETN", 1)
L("IX_OR")
op("PICK", 1); op("CALL", 0, "TDOP"); op("FOLD", 3); op("RETN", 1)
L("IX_XOR")
op("PICK", 1); op("CALL", 0, "TDOP"); op("FOLD", 4); op("RETN", 1)
L("IX_NOT")                                      # mod.rs:1399-1427
op("ONEOF", KL("NOT_INFIX"), "NI_IN", "NI_LIKE", "NI_ILIKE", "NI_BETWEEN", "NI_EXISTS")
for lab, bop in (("NI_IN", "NotIn"), ("NI_LIKE", "NotLike"), ("NI_ILIKE", "NotILike")):
    L(lab)
    op("PUSHI", P["Comparison"]); op("CALL", 0, "TDOP"); op("NODE", NK("BINARY"), B[bop] | (2 << 8)); op("RETN", 1)
L("NI_BETWEEN")
op("PUSHI", P["Between"]); op("CALL", 0, "TDOP"); op("KW", KW["AND"]); op("PUSHI", P["Between"]); op("CALL", 0, "TDOP")
op("NODE", NK("FNCALL"), F["NotBetween"] | (3 << 8)); op("RETN", 1)
L("NI_EXISTS")                                   # `not exists`: needs arguments, error at the NOT token;
                                                 # the reference DROPS `left` (mod.rs:1413-1424: only `args` is kept)
op("IFTT", TT("LParen"), "NI_EXISTS_FAIL")
op("MARK"); op("CALL", 0, "FNARGS_REST"); op("DROPLEFT"); op("CLOSE", NK("FNCALL"), F["NotExists"]); op("RETN", 1)
L("NI_EXISTS_FAIL")
op("FAILAT", PF("NOT_EXISTS_ARGS"), 2)
L("IX_IS")                                       # mod.rs:1430-1438
op("ONEOF", KL("NOT_NULL"), "IS_NOT", "IS_NULL")
L("IS_NOT")
op("KW", KW["NULL"]); op("FOLD", 7); op("RETN", 1)
L("IS_NULL")
op("FOLD", 6); op("RETN", 1)
L("IX_BETWEEN")                                  # mod.rs:1445-1449
op("PUSHI", P["Between"]); op("CALL", 0, "TDOP"); op("KW", KW["AND"]); op("PUSHI", P["Between"]); op("CALL", 0, "TDOP")
op("NODE", NK("FNCALL"), F["Between"] | (3 << 8)); op("RETN", 1)

# ---------------------------------------------------------------------------------------------
# resolve + emit
# ---------------------------------------------------------------------------------------------
prog = []
for w in code:
    if isinstance(w, str):
        assert w in labels, "undefined label " + w
        prog.append(labels[w])
    else:
        prog.append(w)
assert all(0 <= w < 65536 for w in prog)

kl_names = sorted(KWLISTS, key=lambda n: KL(n))
kl_count = max(KL(n) for n in kl_names) + 1
kw_flat, kl_off, kl_len = [], [0] * kl_count, [0] * kl_count
for nme in kl_names:
    kl_off[KL(nme)] = len(kw_flat)
    kl_len[KL(nme)] = len(KWLISTS[nme])
    kw_flat.extend(KW[w.upper()] for w in KWLISTS[nme])
el_count = max(EL(n) for n in EXPECTED) + 1
masks = [0] * el_count
el_lists = [[] for _ in range(el_count)]
for nme, tts in EXPECTED.items():
    for t in tts:
        masks[EL(nme)] |= 1 << TT(t)
    el_lists[EL(nme)] = [TT(t) for t in tts]

out = []
out.append("// GENERATED by gen_parse_program.py -- do not edit.  Included inside namespace npar.")
out.append("enum Op : uint32_t {")
for n, _ in OPS:
    out.append("  OP_%s = %d," % (n, OPNUM[n]))
out.append("};")
out.append("enum Kw : uint32_t {")
for w, i in KW.items():
    out.append("  KW_%s = %d," % (w, i))
out.append("};")
out.append("#define NUTDB_PROGRAM_LEN %d" % len(prog))
out.append("#define NUTDB_PROGRAM_ENTRY %d" % labels["STMT"])
out.append("#define NUTDB_KWLIST_TOTAL %d" % len(kw_flat))
out.append("struct ParseTables {")
out.append("  uint64_t expected_mask[%d];" % el_count)
out.append("  uint16_t program[NUTDB_PROGRAM_LEN];")
out.append("  uint8_t kwlist_off[%d];" % kl_count)
out.append("  uint8_t kwlist_len[%d];" % kl_count)
out.append("  uint8_t kwlist[NUTDB_KWLIST_TOTAL];")
out.append("  uint8_t expected_len[%d];" % el_count)
out.append("  uint8_t expected_list[%d][16];" % el_count)
out.append("};")
out.append("static const ParseTables PARSE_TABLES = {")
out.append("  {" + ", ".join("0x%xull" % m for m in masks) + "},")
out.append("  {")
for i in range(0, len(prog), 20):
    out.append("    " + ", ".join(str(w) for w in prog[i:i + 20]) + ",")
out.append("  },")
out.append("  {" + ", ".join(map(str, kl_off)) + "},")
out.append("  {" + ", ".join(map(str, kl_len)) + "},")
out.append("  {" + ", ".join(map(str, kw_flat)) + "},")
out.append("  {" + ", ".join(str(len(l)) for l in el_lists) + "},")
out.append("  {" + ", ".join("{" + ", ".join(map(str, l + [0] * (16 - len(l)))) + "}" for l in el_lists) + "},")
out.append("};")
open(os.path.join(HERE, "parse_program.h"), "w").write("\n".join(out) + "\n")
print("program words:", len(prog), "labels:", len(labels))
