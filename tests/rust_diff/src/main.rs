//! UNCOMPILED (no Rust toolchain in the build image).
//!
//! Pins the oracle to the Rust reference: for every line of tests/golden/oracle_outputs.tsv
//!     esc(sql) <TAB> OK|ERR <TAB> esc(text)
//! runs nutdb::parser::Parser::parse(sql) and compares format!("{:?}", statement) (OK) or format!("{}", error) (ERR)
//! with the text the C++ oracle (oracle/, the checker of every GPU parity test) produced for the same input.
//! Exit code 0 = the oracle's ASTs, constant folding and error messages / positions are the reference's.
use nutdb::parser::Parser;

fn unesc(t: &str) -> String {
    let mut out = String::with_capacity(t.len());
    let mut it = t.chars();
    while let Some(c) = it.next() {
        if c != '\\' {
            out.push(c);
            continue;
        }
        match it.next() {
            Some('t') => out.push('\t'),
            Some('n') => out.push('\n'),
            Some('r') => out.push('\r'),
            Some(o) => out.push(o), // "\\\\" -> '\\'
            None => {}
        }
    }
    out
}

fn main() {
    let path = std::env::args().nth(1).unwrap_or_else(|| "tests/golden/oracle_outputs.tsv".to_string());
    let data = std::fs::read_to_string(&path).expect("cannot read the golden file");
    let (mut n, mut bad) = (0usize, 0usize);
    for line in data.split('\n').filter(|l| !l.is_empty()) {
        let mut f = line.splitn(3, '\t');
        let (sql, kind, want) = (unesc(f.next().unwrap()), f.next().unwrap(), unesc(f.next().unwrap()));
        // (inputs on which the reference itself panics -- literal.rs:63 -- are not in the file)
        let got = std::panic::catch_unwind(|| match Parser::parse(&sql) {
            Ok(st) => ("OK", format!("{:?}", st)),
            Err(e) => ("ERR", format!("{}", e)),
        });
        n += 1;
        match got {
            Ok((k, text)) if k == kind && text == want => {}
            Ok((k, text)) => {
                bad += 1;
                eprintln!("MISMATCH on {:?}\n  reference: {} {}\n  oracle:    {} {}", sql, k, text, kind, want);
            }
            Err(_) => {
                bad += 1;
                eprintln!("PANIC in the reference on {:?} (oracle: {} {})", sql, kind, want);
            }
        }
    }
    println!("{} statements, {} mismatches", n, bad);
    std::process::exit(if bad == 0 { 0 } else { 1 });
}
