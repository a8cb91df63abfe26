//! See examples/gen_vectors.rs.
