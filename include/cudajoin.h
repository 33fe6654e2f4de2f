/*
 * cudajoin.h — the legacy FFI symbol Kolibrie already binds (library name `cudajoin`,
 * kolibrie/src/cuda/CMakeLists.txt:8, kolibrie/build.rs:75-79).
 *
 * Replaces: kolibrie/src/cuda/cuda_join.cu:48-56 (definition) / kolibrie/src/cuda/cuda_join.rs:14-26 (Rust decl).
 * Caller:   hash_join_cuda (cuda_join.rs:28-60) <- SparqlDatabase::perform_hash_join_cuda_wrapper
 *           (kolibrie/src/sparql_database.rs:3193-3353).
 *
 * Contract kept: same symbol, same signature, `*h_indices` allocated with libc malloc (Rust adopts it with
 * Vec::from_raw_parts, cuda_join.rs:55), `*h_result_count` = number of indices.
 * Contract repaired: every triple is examined (the reference clamps the grid without a
 * grid-stride loop, cuda_join.cu:81-88, and silently drops triples beyond ~300 K); `literal_filter` is honoured
 * (object == *literal_filter; the reference accepts and ignores it, cuda_join.cu:54 vs :91-98); indices are
 * returned ASCENDING (the reference's order is atomicAdd arrival order); CUDA errors yield count 0 and malloc(0).
 */
#ifndef KOLIBRIE_CUDAJOIN_H
#define KOLIBRIE_CUDAJOIN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
__attribute__((visibility("default"))) void perform_hash_join_cuda(
    const uint32_t* h_subjects, const uint32_t* h_predicates, const uint32_t* h_objects,
    uint32_t num_triples, uint32_t predicate_filter,
    uint32_t* literal_filter, /* NULL = no filter */
    uint32_t** h_indices,     /* out: malloc'd by callee */
    uint32_t* h_result_count  /* out */);
#ifdef __cplusplus
}
#endif
#endif
