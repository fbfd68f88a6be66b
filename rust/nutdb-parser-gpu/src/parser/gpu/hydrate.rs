//! UNCOMPILED (no Rust toolchain in the build image) -- see rust/README.md in the nutdb_b200 repository.
//!
//! Flat arrays of `libnutdb_gpu.so` -> the reference's own types: `Statement<'a>` (ast/*.rs) borrowing from the input
//! exactly like the reference (`&'a str` slices, `mod.rs:1886`) and `ParseError` (error.rs:8-56).  Follows
//! `nutdb_b200/csrc/hydrate.cpp` function by function -- that file renders the same walk as `{:?}` text and is tested
//! against the oracle; this one builds the values with the derive(Constructor) `::new` functions (ast/query.rs:21,
//! 145-173) and the enum variants.
//!
//! Wire format (include/nutdb_gpu.h, NUTDB_PN_*): one 32-bit word per node, post-order per statement; `expand()` below
//! is `nutdb_batch_expand_nodes` for one statement (kept in Rust so that re-hydration needs no second FFI call).
use std::borrow::Cow;
use std::str::FromStr;

use bigdecimal::BigDecimal;
use nutdb_gpu_sys as sys;
use nutdb_gpu_sys::*; // NK_* / PN_* / ST_* constants

use crate::parser::literal::{unescape_double_quoted_string, unescape_single_quoted_string};
use crate::parser::tokenizer::{Position, TokenType, TokenizeError, TokenizeErrorType};
use crate::parser::*;

/// An expanded node of one statement (the fields of `NutdbNode` that re-hydration uses).
#[derive(Clone, Copy, Default)]
struct Node {
    kind: u8,
    sub: u8,
    aux: u16,
    a: u32, // leaf: span start; interior: first node of the subtree
    b: u32, // leaf: span end
}

fn find_ext<'b>(ext: &'b [sys::NutdbNodeExt], index: u32) -> Option<&'b sys::NutdbNodeExt> {
    ext.binary_search_by_key(&index, |e| e.index).ok().map(|i| &ext[i])
}

/// Wire words of one statement -> nodes (nutdb_batch_expand_nodes, hydrate.cpp `expand_stmt_nodes`).
fn expand(words: &[u32], first_index: u32, ext: &[sys::NutdbNodeExt]) -> Vec<Node> {
    let mut out = Vec::with_capacity(words.len());
    let mut pos = 0u32;
    for (j, &x) in words.iter().enumerate() {
        let mut n = Node { kind: (x & 127) as u8, sub: ((x >> PN_SUB_SHIFT) & 31) as u8, aux: ((x >> PN_FLAG_SHIFT) & 1) as u16, a: 0, b: 0 };
        if n.kind >= NK_FIRST_INTERIOR {
            let size = x >> PN_SIZE_SHIFT;
            if size == PN_SIZE_EXT {
                let e = find_ext(ext, first_index + j as u32).expect("side-table entry of an escaped node");
                n.kind = (e.hdr & 255) as u8;
                n.sub = ((e.hdr >> 8) & 255) as u8;
                n.aux = (e.hdr >> 16) as u16;
                n.a = e.a;
            } else {
                n.a = j as u32 - size;
            }
        } else {
            let gap = (x >> PN_GAP_SHIFT) & 1023;
            let len = x >> PN_LEN_SHIFT;
            if len == PN_LEN_SPECIAL && gap == PN_GAP_NOSPAN {
                // no source span
            } else if len == PN_LEN_SPECIAL && gap == PN_GAP_EXT {
                let e = find_ext(ext, first_index + j as u32).expect("side-table entry of an escaped node");
                n.kind = (e.hdr & 255) as u8;
                n.sub = ((e.hdr >> 8) & 255) as u8;
                n.aux = ((e.hdr >> 16) & 1) as u16;
                n.a = e.a;
                n.b = e.a + e.b;
                pos = n.b;
            } else {
                n.a = pos + gap;
                n.b = n.a + len;
                pos = n.b;
            }
        }
        out.push(n);
    }
    out
}

/// Re-hydrates every statement of a batch.  `sqls[i]` is statement i's own text.
///
/// # Safety
/// `b` must be a live batch of the library (or a chunk handed to a dispatcher callback) with host-resident arrays.
pub(super) unsafe fn batch<'a>(b: &sys::NutdbBatch, sqls: &[&'a str]) -> Vec<Result<Statement<'a>, ParseError>> {
    // statement records: the 24-byte form, or the 8-byte wire form (F_WIRE_STMT) expanded here
    let expanded: Vec<sys::NutdbStmt>;
    let stmts: &[sys::NutdbStmt] = if !b.stmt.is_null() || b.n_stmt == 0 {
        if b.n_stmt == 0 { &[] } else { core::slice::from_raw_parts(b.stmt, b.n_stmt as usize) }
    } else {
        let w = core::slice::from_raw_parts(b.wstmt, b.n_stmt as usize);
        let mut node_begin = 0u32;
        expanded = w.iter().map(|&x| {
            let s = sys::NutdbStmt { status: (x & 15) as u32, tok_begin: 0, tok_count: 0, node_begin,
                                     node_count: ((x >> 4) & 0x3FFF_FFFF) as u32, tok_used: (x >> 34) as u32 };
            node_begin += s.node_count;
            s
        }).collect();
        &expanded
    };
    let words = if b.n_node > 0 { core::slice::from_raw_parts(b.pnode, b.n_node as usize) } else { &[][..] };
    let errs = if b.n_err > 0 { core::slice::from_raw_parts(b.err, b.n_err as usize) } else { &[][..] };
    let ext = if b.n_ext > 0 { core::slice::from_raw_parts(b.ext, b.n_ext as usize) } else { &[][..] };
    let mut next_err = 0usize; // error records are sorted by statement
    let mut out = Vec::with_capacity(stmts.len());
    for (i, s) in stmts.iter().enumerate() {
        if s.status == ST_OK {
            let w = &words[s.node_begin as usize..(s.node_begin + s.node_count) as usize];
            let nodes = expand(w, s.node_begin, ext);
            let h = H { nd: &nodes, sql: sqls[i] };
            out.push(Ok(h.statement(nodes.len() - 1)));
        } else {
            while errs[next_err].stmt < i as u32 {
                next_err += 1;
            }
            out.push(Err(error(&errs[next_err], sqls[i])));
        }
    }
    out
}

struct H<'n, 'a> {
    nd: &'n [Node],
    sql: &'a str,
}

fn interior(n: &Node) -> bool {
    n.kind >= NK_FIRST_INTERIOR
}

impl<'n, 'a> H<'n, 'a> {
    fn span(&self, i: usize) -> &'a str {
        &self.sql[self.nd[i].a as usize..self.nd[i].b as usize]
    }
    /// first node of the subtree rooted at r (a qualified identifier owns the QUAL leaf in front of it)
    fn subtree_start(&self, r: usize) -> usize {
        let n = &self.nd[r];
        if interior(n) {
            n.a as usize
        } else if n.kind == NK_IDENT && (n.aux & 1) != 0 && r > 0 {
            r - 1
        } else {
            r
        }
    }
    /// children of interior node i, left to right
    fn kids(&self, i: usize) -> Vec<usize> {
        let first = self.nd[i].a as usize;
        let mut k = Vec::new();
        let mut r = i as isize - 1;
        while r >= first as isize {
            k.push(r as usize);
            r = self.subtree_start(r as usize) as isize - 1;
        }
        k.reverse();
        k
    }
    fn node_int(&self, i: usize) -> u128 {
        // digits were validated on the device (integer_from_str!, literal.rs:18-31)
        let radix = if (self.nd[i].aux & 1) != 0 { 16 } else { 10 };
        u128::from_str_radix(self.span(i), radix).unwrap_or(0)
    }
    fn str_value(&self, i: usize) -> Cow<'a, str> {
        let s = self.span(i);
        match self.nd[i].sub {
            0 => Cow::Borrowed(s),                                                    // raw string literal
            1 => Cow::Owned(unescape_single_quoted_string(s).expect("validated on the device")),
            _ => Cow::Owned(unescape_double_quoted_string(s).expect("validated on the device")),
        }
    }

    // ------------------------------------------------------------------ expressions (hydrate.cpp: identifier / fncall / expr)
    fn identifier(&self, i: usize) -> Identifier<'a> {
        let name = if self.nd[i].sub == 1 { IdentifierName::Wildcard } else { IdentifierName::Word(self.span(i)) };
        let qualifier = if (self.nd[i].aux & 1) != 0 { Some(self.span(i - 1)) } else { None };
        Identifier::new(name, qualifier)
    }
    fn fncall(&self, i: usize) -> FnCall<'a> {
        let k = self.kids(i);
        let (callee, first) = match self.nd[i].sub {
            0 => (FnName::If, 0),
            1 => (FnName::MultiIf, 0),
            2 => (FnName::CaseWhen, 0),
            3 => (FnName::Between, 0),
            4 => (FnName::NotBetween, 0),
            5 => (FnName::Exists, 0),
            6 => (FnName::NotExists, 0),
            _ => (FnName::Others(self.span(k[0])), 1),
        };
        FnCall::new(callee, k[first..].iter().map(|&c| self.expr(c)).collect())
    }
    fn expr(&self, i: usize) -> Expr<'a> {
        let x = &self.nd[i];
        match x.kind {
            NK_IDENT => Expr::Identifier(self.identifier(i)),
            NK_QPARAM => Expr::QueryParameter(QueryParameter::new(self.node_int(i) as usize)),
            NK_LIT_INT => Expr::Literal(Literal::Integer(self.node_int(i), (x.sub & 1) == 0)),
            NK_LIT_FLOAT => {
                let d = BigDecimal::from_str(self.span(i)).expect("the lexer only produces digits '.' digits");
                Expr::Literal(Literal::Float(Box::new(if (x.sub & 1) != 0 { -d } else { d })))
            }
            NK_LIT_STR => Expr::Literal(Literal::String(self.str_value(i))),
            NK_LIT_BOOL => Expr::Literal(Literal::Boolean(x.sub != 0)),
            NK_LIT_NULL => Expr::Literal(Literal::Null),
            NK_LIT_INTERVAL => {
                let unit = match x.sub {
                    0 => IntervalUnit::Second,
                    1 => IntervalUnit::Minute,
                    2 => IntervalUnit::Hour,
                    3 => IntervalUnit::Day,
                    4 => IntervalUnit::Month,
                    _ => IntervalUnit::Year,
                };
                Expr::Literal(Literal::Interval(self.node_int(i) as u64, unit))
            }
            NK_COLLECTION => {
                let typ = match x.sub {
                    0 => CollectionType::Tuple,
                    1 => CollectionType::Map,
                    _ => CollectionType::Array,
                };
                Expr::Collection(Collection::new(typ, self.kids(i).iter().map(|&c| self.expr(c)).collect()))
            }
            NK_UNARY => {
                let op = match x.sub {
                    0 => UnaryOperator::BitwiseNot,
                    1 => UnaryOperator::Not,
                    2 => UnaryOperator::IsNull,
                    _ => UnaryOperator::IsNotNull,
                };
                Expr::UnaryOp(UnaryOp::new(op, Box::new(self.expr(self.kids(i)[0]))))
            }
            NK_BINARY => {
                use BinaryOperator::*;
                const OPS: [BinaryOperator; 26] = [Plus, Minus, Multi, Div, Mod, Gt, Lt, GtEq, LtEq, Eq, NotEq, And, Or, Xor, Like,
                                                   NotLike, ILike, NotILike, In, NotIn, IndexAccess, BitwiseOr, BitwiseAnd,
                                                   BitwiseXor, BitwiseLeftShift, BitwiseRightShift];
                let k = self.kids(i);
                Expr::BinaryOp(BinaryOp::new(OPS[x.sub as usize], Box::new(self.expr(k[0])), Box::new(self.expr(k[1]))))
            }
            NK_FNCALL => Expr::FnCall(self.fncall(i)),
            NK_QUERY_BODY | NK_QUERY_UNION => Expr::Subquery(self.query(i)),
            _ => unreachable!("node kind {} is not an expression", x.kind),
        }
    }

    // ------------------------------------------------------------------ queries (hydrate.cpp: query_expr / source / query / body)
    /// (expr ALIAS?) starting at k[*j]; consumes one QueryExpr
    fn query_expr(&self, k: &[usize], j: &mut usize) -> QueryExpr<'a> {
        let inner = self.expr(k[*j]);
        *j += 1;
        let alias = if *j < k.len() && self.nd[k[*j]].kind == NK_ALIAS {
            *j += 1;
            Some(self.span(k[*j - 1]))
        } else {
            None
        };
        QueryExpr::new(inner, alias)
    }
    fn query_expr_list(&self, i: usize) -> Vec<QueryExpr<'a>> {
        let k = self.kids(i);
        let mut j = 0;
        let mut v = Vec::new();
        while j < k.len() {
            v.push(self.query_expr(&k, &mut j));
        }
        v
    }
    fn source(&self, k: &[usize], j: &mut usize) -> QuerySource<'a> {
        let s = k[*j];
        *j += 1;
        let inner = match self.nd[s].kind {
            NK_FNCALL => DataSource::TableFn(self.fncall(s)),
            NK_IDENT => DataSource::Table(self.span(s)),
            _ => DataSource::Subquery(self.query(s)),
        };
        let alias = if *j < k.len() && self.nd[k[*j]].kind == NK_ALIAS {
            *j += 1;
            Some(self.span(k[*j - 1]))
        } else {
            None
        };
        QuerySource::new(inner, alias)
    }
    fn query(&self, i: usize) -> Query<'a> {
        if self.nd[i].kind == NK_QUERY_BODY {
            return Query::Single(Box::new(self.body(i)));
        }
        let k = self.kids(i);
        let typ = match self.nd[i].sub {
            0 => UnionType::UnionAll,
            1 => UnionType::UnionDistinct,
            2 => UnionType::Intersect,
            _ => UnionType::Except,
        };
        Query::Union { typ, left: Box::new(self.query(k[0])), right: Box::new(self.query(k[1])) }
    }
    fn body(&self, i: usize) -> QueryBody<'a> {
        let k = self.kids(i);
        let mut j = 0usize;
        let is = |j: usize, kind: u8| j < k.len() && self.nd[k[j]].kind == kind;
        let with = if is(j, NK_WITH) {
            let w = self.kids(k[j]);
            j += 1;
            let mut ctes = Vec::new();
            let mut c = 0;
            while c + 1 < w.len() {
                ctes.push(QueryCTE::new(Box::new(self.query(w[c + 1])), self.span(w[c])));
                c += 2;
            }
            Some(WithClause::new(ctes))
        } else {
            None
        };
        let distinct = if is(j, NK_DISTINCT) {
            let d = k[j];
            j += 1;
            Some(DistinctClause::new(if (self.nd[d].aux & 1) != 0 { Some(self.query_expr_list(d)) } else { None }))
        } else {
            None
        };
        let columns = self.query_expr_list(k[j]);
        j += 1;
        let from = if is(j, NK_FROM) {
            let f = self.kids(k[j]);
            j += 1;
            let mut p = 0;
            Some(FromClause::new(self.source(&f, &mut p)))
        } else {
            None
        };
        let mut joins = Vec::new();
        while is(j, NK_JOIN) {
            let jn = k[j];
            j += 1;
            let c = self.kids(jn);
            let mut p = 0;
            let typ = match self.nd[jn].sub {
                0 => JoinType::Inner,
                1 => JoinType::FullOuter,
                2 => JoinType::LeftOuter,
                3 => JoinType::RightOuter,
                4 => JoinType::LeftSemi,
                5 => JoinType::RightSemi,
                6 => JoinType::LeftAnti,
                _ => JoinType::RightAnti,
            };
            let source = self.source(&c, &mut p);
            let condition = if (self.nd[jn].aux & 1) != 0 {
                JoinCondition::Using(c[p..].iter().map(|&q| self.identifier(q)).collect())
            } else {
                JoinCondition::On(Box::new(self.expr(c[p])))
            };
            joins.push(JoinClause::new(typ, source, condition));
        }
        let r#where = if is(j, NK_WHERE) {
            j += 1;
            Some(WhereClause::new(self.expr(self.kids(k[j - 1])[0])))
        } else {
            None
        };
        let group_by = if is(j, NK_GROUPBY) {
            j += 1;
            Some(GroupByClause::new(self.query_expr_list(k[j - 1])))
        } else {
            None
        };
        let having = if is(j, NK_HAVING) {
            j += 1;
            Some(HavingClause::new(self.expr(self.kids(k[j - 1])[0])))
        } else {
            None
        };
        let order_by = if is(j, NK_ORDERBY) {
            let c = self.kids(k[j]);
            j += 1;
            let mut p = 0;
            let mut keys = Vec::new();
            while p < c.len() {
                let e = self.query_expr(&c, &mut p);
                let dir = if p < c.len() && self.nd[c[p]].kind == NK_ORDER_DESC {
                    p += 1;
                    OrderDirection::DESC
                } else {
                    OrderDirection::ASC
                };
                keys.push(QueryOrderKey::new(e, dir));
            }
            Some(OrderByClause::new(keys))
        } else {
            None
        };
        let limit = if is(j, NK_LIMIT) {
            let l = k[j];
            let c = self.kids(l);
            let first = self.node_int(c[0]) as usize;
            let second = if c.len() > 1 { self.node_int(c[1]) as usize } else { 0 };
            // sub 0: LIMIT n; 1: LIMIT offset, n; 2: LIMIT n OFFSET offset (mod.rs:503-544)
            let (size, offset) = match self.nd[l].sub {
                0 => (first, 0),
                1 => (second, first),
                _ => (first, second),
            };
            Some(LimitClause::new(size, offset, (self.nd[l].aux & 1) != 0))
        } else {
            None
        };
        QueryBody::new(with, distinct, columns, from, joins, r#where, group_by, having, order_by, limit)
    }

    // ------------------------------------------------------------------ DDL (hydrate.cpp: datatype / coldef / condef / idxdef)
    fn datatype(&self, i: usize) -> DataType<'a> {
        use ScalarDataType::*;
        let x = &self.nd[i];
        let scalar = |sub: u8, param: usize| -> ScalarDataType {
            match sub {
                0 => Int8, 1 => Int16, 2 => Int32, 3 => Int64, 4 => Int128, 5 => UInt8, 6 => UInt16, 7 => UInt32, 8 => UInt64,
                9 => UInt128, 10 => Serial32, 11 => Serial64, 12 => Serial128, 13 => USerial32, 14 => USerial64, 15 => USerial128,
                16 => Decimal32 { scale: param as u8 }, 17 => Decimal64 { scale: param as u8 }, 18 => Float32, 19 => Float64,
                20 => Boolean, 21 => Chars { length: param }, 22 => String { max_length: param }, 23 => Uuid, 24 => Date,
                _ => Datetime,
            }
        };
        if x.kind == NK_DT_SCALAR {
            return DataType::Scalar(scalar(x.sub, 0)); // plain `String` = String { max_length: 0 }
        }
        let k = self.kids(i);
        if x.kind == NK_DT_PARAM {
            return DataType::Scalar(scalar(x.sub, self.node_int(k[0]) as usize));
        }
        DataType::Compound(match x.sub {
            0 => CompoundDataType::Array(Box::new(self.datatype(k[0]))),
            1 => {
                // Enum('a' = 1, 'b', ..): an explicit `= n` restarts the counter (mod.rs:1799-1813)
                let mut binds = Vec::new();
                let mut id = 0usize;
                let mut j = 0;
                while j < k.len() {
                    let lit = k[j];
                    j += 1;
                    if j < k.len() && self.nd[k[j]].kind == NK_NUM {
                        id = self.node_int(k[j]) as usize;
                        j += 1;
                    }
                    binds.push(EnumBind::new(id, self.str_value(lit)));
                    id += 1;
                }
                CompoundDataType::Enum(binds)
            }
            2 => CompoundDataType::Tuple(k.iter().map(|&c| self.datatype(c)).collect()),
            // Map(K, V) is stored as Map(Box(V), Box(K)) (mod.rs:1776-1780)
            3 => CompoundDataType::Map(Box::new(self.datatype(k[1])), Box::new(self.datatype(k[0]))),
            4 => CompoundDataType::Dictionary(Box::new(self.datatype(k[0]))),
            _ => CompoundDataType::Nullable(Box::new(self.datatype(k[0]))),
        })
    }
    fn expr_list_of(&self, i: usize) -> Vec<Expr<'a>> {
        self.kids(i).iter().map(|&c| self.expr(c)).collect()
    }
    fn coldef(&self, i: usize) -> ColumnDefinition<'a> {
        let k = self.kids(i);
        let mut default = None;
        let mut comment = None;
        for &c in &k[2..] {
            if self.nd[c].kind == NK_ATTR_DEFAULT {
                default = Some(self.expr(self.kids(c)[0]));
            } else {
                comment = Some(self.str_value(c));
            }
        }
        ColumnDefinition::new(self.span(k[0]), self.datatype(k[1]), default, comment)
    }
    fn condef(&self, i: usize) -> ConstraintDefinition<'a> {
        let k = self.kids(i);
        ConstraintDefinition::new(self.span(k[0]), self.expr(k[1]))
    }
    fn idxdef(&self, i: usize) -> IndexDefinition<'a> {
        let k = self.kids(i);
        IndexDefinition::new(self.span(k[0]), self.fncall(k[1]))
    }

    // ------------------------------------------------------------------ statements (hydrate.cpp: statement)
    fn statement(&self, i: usize) -> Statement<'a> {
        let x = &self.nd[i];
        let k = self.kids(i);
        let flag = (x.aux & 1) != 0;
        match x.kind {
            NK_STMT_SELECT => Statement::Select(SelectStmt::new(self.query(k[0]))),
            NK_STMT_EXPLAIN => Statement::Explain(ExplainStmt::new(self.query(k[0]))),
            NK_STMT_INSERT => {
                let last = k.len() - 1;
                let column_list = if last > 1 { Some(k[1..last].iter().map(|&c| self.span(c)).collect()) } else { None };
                let d = k[last];
                let data = match self.nd[d].kind {
                    NK_ROWS => {
                        let rows = self.kids(d);
                        let column_size = self.kids(rows[0]).len();
                        let mut data = Vec::new();
                        for &r in &rows {
                            data.extend(self.kids(r).iter().map(|&e| self.expr(e)));
                        }
                        InsertSource::Rows { column_size, data }
                    }
                    NK_FNCALL => InsertSource::FnCall(self.fncall(d)),
                    _ => InsertSource::Subquery(self.query(d)),
                };
                Statement::Insert(InsertStmt::new(self.span(k[0]), column_list, data))
            }
            NK_STMT_ALTER => {
                let e = k[1];
                let action = match x.sub {
                    0 => AlterAction::Add {
                        entity: match self.nd[e].kind {
                            NK_COLDEF => AddableEntity::Column(self.coldef(e)),
                            NK_INDEXDEF => AddableEntity::Index(self.idxdef(e)),
                            _ => AddableEntity::Constraint(self.condef(e)),
                        },
                        if_not_exists: flag,
                        position: if k.len() > 2 && self.nd[k[2]].kind == NK_POS_FIRST {
                            EntityPosition::First
                        } else if k.len() > 2 && self.nd[k[2]].kind == NK_POS_AFTER {
                            EntityPosition::After(self.span(k[2]))
                        } else {
                            EntityPosition::Last
                        },
                    },
                    1 => AlterAction::Drop {
                        entity: if self.nd[e].kind == NK_STR {
                            DroppableEntity::Partition(self.str_value(e))
                        } else {
                            match self.nd[e].sub {
                                0 => DroppableEntity::Column(self.span(e)),
                                1 => DroppableEntity::Index(self.span(e)),
                                _ => DroppableEntity::Constraint(self.span(e)),
                            }
                        },
                        if_exists: flag,
                    },
                    _ => AlterAction::Rename {
                        entity: match self.nd[e].sub {
                            0 => RenamableEntity::Column(self.span(e)),
                            1 => RenamableEntity::Index(self.span(e)),
                            2 => RenamableEntity::Constraint(self.span(e)),
                            _ => RenamableEntity::Table,
                        },
                        new_name: self.span(k[2]),
                    },
                };
                Statement::Alter(AlterStmt::new(Alter::new(action, self.span(k[0]))))
            }
            NK_STMT_CREATE => {
                let d = k[0];
                let c = self.kids(d);
                // primary_key / order_by / partition_by / comment / strategy wherever they appeared
                let (mut pk, mut ob, mut pb, mut com, mut strat) = (None, None, None, None, None);
                let is_table = self.nd[d].kind == NK_TABLEDEF;
                let attrs = if is_table { &c[1..] } else { &c[1..c.len() - 1] };
                for &q in attrs {
                    match self.nd[q].kind {
                        NK_ATTR_PK => pk = Some(self.expr_list_of(q)),
                        NK_ATTR_ORDER => ob = Some(self.expr_list_of(q)),
                        NK_ATTR_PART => pb = Some(self.expr(self.kids(q)[0])),
                        NK_STR => com = Some(self.str_value(q)),
                        NK_STRATEGY => strat = Some(self.span(q)),
                        _ => {}
                    }
                }
                let entity = if is_table {
                    let pick = |kind: u8| c[1..].iter().copied().filter(|&q| self.nd[q].kind == kind).collect::<Vec<_>>();
                    CreatableEntity::Table(TableDefinition::new(
                        self.span(c[0]),
                        pick(NK_COLDEF).into_iter().map(|q| self.coldef(q)).collect(),
                        pick(NK_CONSTRDEF).into_iter().map(|q| self.condef(q)).collect(),
                        pick(NK_INDEXDEF).into_iter().map(|q| self.idxdef(q)).collect(),
                        pk, ob, pb, com,
                    ))
                } else {
                    CreatableEntity::View(ViewDefinition::new(self.span(c[0]), strat.unwrap_or(""), pk, ob, pb,
                                                              self.query(*c.last().unwrap()), com))
                };
                Statement::Create(CreateStmt::new(flag, entity))
            }
            NK_STMT_DESCRIBE => Statement::Describe(DescribeStmt::new(match x.sub {
                2 => DescribableEntity::Database,
                0 => DescribableEntity::Table(self.span(k[0])),
                _ => DescribableEntity::View(self.span(k[0])),
            })),
            NK_STMT_DROP => Statement::Drop(DropStmt::new(if x.sub == 0 { DatabaseEntity::Table } else { DatabaseEntity::View }, flag,
                                                          self.span(k[0]))),
            NK_STMT_TRUNCATE => Statement::Truncate(TruncateStmt::new(
                if x.sub == 0 { DatabaseEntity::Table } else { DatabaseEntity::View }, flag, self.span(k[0]))),
            NK_STMT_OPTIMIZE => Statement::Optimize(OptimizeStmt::new(self.span(k[0]), if k.len() > 1 { Some(self.expr(k[1])) } else { None })),
            NK_STMT_SET => Statement::Set(SetStmt::new(self.span(k[0]), self.expr(k[1]))),
            _ => unreachable!("node kind {} is not a statement", x.kind),
        }
    }
}

// ---------------------------------------------------------------------- errors (hydrate.cpp: lex_error_text / syntax_error_text)
fn char_at(sql: &str, pos: u32) -> String {
    sql.get(pos as usize..).and_then(|s| s.chars().next()).map(|c| c.to_string()).unwrap_or_default()
}

/// One error record -> the reference's ParseError (error.rs:8-56, tokenizer/error.rs:7-30).
fn error(e: &sys::NutdbError, sql: &str) -> ParseError {
    let pos = Position::new(e.line as usize, e.col as usize);
    if e.cls as u32 == ST_LEX_ERROR {
        use TokenizeErrorType::*;
        let c = format!("'{}'", char_at(sql, e.pos));
        // one site code per emit_error! of tokenizer/mod.rs (NUTDB_LE_*, line numbers in include/nutdb_gpu.h)
        let (t, ctx) = match e.code {
            1 => (UnexpectedChar, format!("{c} is invalid outside string literal")),
            2 => (UnexpectedChar, "\\r in string is supported but should be escaped by '\\'".to_string()),
            3 => (UnexpectedChar, "\\n in string is supported but should be escaped by '\\'".to_string()),
            4 => (UnexpectedEOF, "string literal is not complete".to_string()),
            5 => (UnexpectedChar, format!("{c} is invalid in numeric literal")),
            6 => (UnexpectedChar, format!("{c} cannot be a part of integer literal")),
            7 => (UnexpectedChar, format!("{c} cannot be a part of float literal")),
            8 => (UnexpectedChar, format!("{c} cannot be a part of identifier or keyword")),
            9 => (UnexpectedChar, "config identifier cannot starts with numbers".to_string()),
            10 => (UnexpectedChar, format!("{c} cannot be a part of config identifier")),
            11 => (Incomplete, "identifier should have name".to_string()),
            12 => (Incomplete, "delimited identifier cannot be an empty string".to_string()),
            13 => (UnexpectedChar, "'\\r' or '\\n' cannot be a part of delimited identifier".to_string()),
            14 => (UnexpectedEOF, "delimited identifier is not complete".to_string()),
            15 => (UnexpectedChar, format!("{c} cannot be a part of query parameter")),
            16 => (Incomplete, "query parameter should have an index".to_string()),
            17 => (UnexpectedChar, "'!' can only be used with '='".to_string()),
            _ => (UnexpectedEOF, "block comment is not complete".to_string()),
        };
        return ParseError::LexError(TokenizeError { t, ctx, pos });
    }
    let raw = || sql.get(e.b as usize..e.c as usize).unwrap_or("").to_string();
    ParseError::SyntaxError(match e.code {
        1 => SyntaxError::NotExpectedTokenTypes { expected: expected_types(e.a), actual: token_type(e.b), pos },
        2 => {
            let (expected, actual) = if e.a >= 1000 {
                (vec![KEYWORD_TEXT[(e.a - 1001) as usize].to_string()], raw())
            } else if e.a == 19 {
                (expected_keywords(e.a), "as".to_string()) // mod.rs:826 reports the constant AS
            } else {
                (expected_keywords(e.a), raw())
            };
            SyntaxError::NotExpectedKeywords { expected, actual, pos }
        }
        3 => {
            const MSGS: [&str; 10] = ["?", "statements should start with a keyword", "more than one statement",
                                      "cannot recognize statement", "not a subquery",
                                      "query source must be a subquery, a table function or a table",
                                      "insert source must be a subquery, values, or a function call",
                                      "indexer must be a function call", "`not exists` should have arguments",
                                      "`exists` should have arguments"];
            SyntaxError::ParseFail { msg: MSGS[(e.a as usize).min(9)].to_string(), pos }
        }
        4 => SyntaxError::EmptyQuery,
        5 => SyntaxError::InvalidEscapedUnicode { hex: raw() },
        // the `source` fields are std errors without public constructors: re-run the conversion that failed
        6 => SyntaxError::InvalidFloatLiteral { raw: raw(), source: BigDecimal::from_str(&raw()).unwrap_err() },
        7 => SyntaxError::InvalidHexLiteral { raw: raw(), source: u128::from_str_radix(&raw(), 16).unwrap_err() },
        8 => SyntaxError::InvalidIntegerLiteral { raw: raw(), source: u128::from_str(&raw()).unwrap_err() },
        _ => {
            const WHAT: [&str; 7] = ["?", "primary key", "order by", "partition by", "comment", "update by", "default"];
            let (this, that) = if e.a == 7 {
                (format!("row has {} column(s)", e.b), format!("previous rows have {} column(s)", e.c))
            } else {
                (WHAT[(e.a as usize).min(6)].to_string(), WHAT[(e.a as usize).min(6)].to_string())
            };
            SyntaxError::Conflicts { this, that, pos }
        }
    })
}

// ---- tables generated from include/nutdb_gpu.h, csrc/lex_tables.hpp and csrc/gen_parse_program.py ----
/// keyword id (1-based, keyword.rs order) -> text
static KEYWORD_TEXT: [&str; 115] = ["by", "as", "on", "from", "intersect", "union", "all", "except", "distinct", "with", "select", "join", "where", "group", "having", "order", "limit", "offset", "using", "ties", "asc", "desc", "explain", "insert", "into", "values", "create", "primary", "key", "comment", "update", "default", "check", "describe", "drop", "alter", "add", "rename", "first", "after", "truncate", "optimize", "set", "database", "table", "view", "column", "index", "constraint", "partition", "null", "true", "false", "and", "or", "xor", "not", "in", "exists", "if", "case", "when", "then", "else", "end", "is", "between", "like", "ilike", "interval", "second", "minute", "hour", "day", "month", "year", "int8", "int16", "int32", "int64", "int128", "uint8", "uint16", "uint32", "uint64", "uint128", "serial32", "serial64", "serial128", "userial32", "userial64", "userial128", "decimal32", "decimal64", "float32", "float64", "boolean", "chars", "string", "uuid", "date", "datetime", "array", "enum", "tuple", "map", "dictionary", "nullable", "inner", "outer", "left", "right", "full", "semi", "anti"];
/// NUTDB_TT_* ordinal -> TokenType (token.rs:5-87)
fn token_type(t: u32) -> TokenType {
    use TokenType::*;
    match t {
        0 => KeywordOrIdentifier,
        1 => DelimitedIdentifier,
        2 => ConfigIdentifier,
        3 => QueryParameter,
        4 => RawStringLiteral,
        5 => EscapedSQStringLiteral,
        6 => EscapedDQStringLiteral,
        7 => IntegerLiteral,
        8 => FloatLiteral,
        9 => HexLiteral,
        10 => Comma,
        11 => Dot,
        12 => Colon,
        13 => SemiColon,
        14 => Plus,
        15 => Minus,
        16 => Mul,
        17 => Div,
        18 => Mod,
        19 => Eq,
        20 => NotEq,
        21 => Lt,
        22 => Gt,
        23 => LtEq,
        24 => GtEq,
        25 => LParen,
        26 => RParen,
        27 => LBracket,
        28 => RBracket,
        29 => LBrace,
        30 => RBrace,
        31 => BitAnd,
        32 => BitOr,
        33 => BitXor,
        34 => BitNot,
        35 => BitLShift,
        36 => BitRShift,
        37 => Comment,
        38 => Whitespace,
        39 => EOF,
        _ => EOF,
    }
}
/// NUTDB_EL_* -> the `expected` list of SyntaxError::NotExpectedTokenTypes
fn expected_types(id: u32) -> Vec<TokenType> {
    use TokenType::*;
    match id {
        1 => vec![RParen],  // RParen
        2 => vec![LParen],  // LParen
        3 => vec![RBracket],  // RBracket
        4 => vec![RBrace],  // RBrace
        5 => vec![IntegerLiteral, HexLiteral, FloatLiteral],  // NegLiteral
        6 => vec![DelimitedIdentifier, KeywordOrIdentifier, Mul],  // Identifier
        7 => vec![Colon],  // Colon
        8 => vec![KeywordOrIdentifier],  // Keyword
        9 => vec![KeywordOrIdentifier, DelimitedIdentifier],  // IdentString
        10 => vec![IntegerLiteral, HexLiteral],  // IntLiteral
        11 => vec![RawStringLiteral, EscapedSQStringLiteral, EscapedDQStringLiteral],  // StrLiteral
        12 => vec![ConfigIdentifier],  // ConfigIdent
        13 => vec![Eq],  // Eq
        14 => vec![Comma],  // Comma
        15 => vec![RawStringLiteral, EscapedSQStringLiteral, EscapedDQStringLiteral, FloatLiteral, HexLiteral, IntegerLiteral, QueryParameter, KeywordOrIdentifier, DelimitedIdentifier, LParen, LBracket, LBrace, Minus, Plus, BitNot, Mul],  // Prefix
        _ => vec![],
    }
}
/// NUTDB_KL_* -> the `expected` list of SyntaxError::NotExpectedKeywords
fn expected_keywords(id: u32) -> Vec<String> {
    let l: &[&str] = match id {
        1 => &["with", "select"],  // WITH_SELECT
        2 => &["all", "distinct"],  // ALL_DISTINCT
        3 => &["on", "using"],  // ON_USING
        4 => &["values", "from", "select", "with"],  // INSERT_SOURCE
        5 => &["table", "view"],  // TABLE_VIEW
        6 => &["primary", "order", "partition", "comment"],  // TABLE_ATTRS
        7 => &["as", "update", "primary", "order", "partition", "comment"],  // VIEW_ATTRS
        8 => &["default", "comment"],  // COLUMN_ATTRS
        9 => &["add", "drop", "rename"],  // ALTER_ACTION
        10 => &["column", "index", "constraint"],  // ADD_ENTITY
        11 => &["column", "index", "constraint", "partition"],  // DROP_ENTITY
        12 => &["column", "index", "constraint", "table"],  // RENAME_ENTITY
        13 => &["table", "view", "database"],  // DESCRIBE_ENTITY
        14 => &["in", "like", "ilike", "between", "exists"],  // NOT_INFIX
        15 => &["not", "null"],  // NOT_NULL
        16 => &["second", "minute", "hour", "day", "month", "year"],  // INTERVAL_UNIT
        17 => &["int8", "int16", "int32", "int64", "int128", "uint8", "uint16", "uint32", "uint64", "uint128", "serial32", "serial64", "serial128", "userial32", "userial64", "userial128", "decimal32", "decimal64", "float32", "float64", "boolean", "chars", "string", "uuid", "date", "datetime", "array", "enum", "tuple", "map", "dictionary", "nullable"],  // DATATYPE
        18 => &["when", "else", "end"],  // CASE_NEXT
        19 => &["update"],  // VIEW_NEEDS_UPDATE
        _ => &[],
    };
    l.iter().map(|s| s.to_string()).collect()
}
