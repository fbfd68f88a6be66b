"""nutdb_b200 -- B200-native (sm_100a) SQL lexer/parser behind the C ABI in include/nutdb_gpu.h.

Drop-in for the one hot path of nutdb/nutdb: `nutdb::parser::Parser::parse` (reference
src/parser/mod.rs:26-29).  `gpu` binds the CUDA library, `workload` generates benchmark batches.
"""
__all__ = ["gpu", "workload", "build", "parser", "dispatch", "stream"]
