/*
 * rpt_oracle.h -- CPU restatement of the ocramz/rp-tree hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a plain-C restatement of the reference's Haskell algorithm for
 *   forestBatch / forest (single chunk and multi-chunk)  ->  candidates / knn / knnPQ  ->  recallWith.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (rp-tree_b200/) never links, imports or calls anything here.
 *
 * PARITY STATUS: "parity unpinned" for everything except (i) the four vector-space known-answer
 * tests of test/Data/RPTreeSpec.hs:22-46 and (ii) the SplitMix64 core (haddock vector mkSMGen 42).
 * The reference cannot be compiled here (no GHC/cabal/stack) and its own tests seed from system
 * entropy (test/Data/RPTreeSpec.hs:49), so no golden tree / leaf set / knn output exists anywhere.
 * Each function cites the reference file:line it restates.
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math   (no FMA contraction: GHC emits separate mul/add)
 */
#ifndef RPT_ORACLE_H
#define RPT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- SplitMix64 (Hackage splitmix-0.1.0.3, System.Random.SplitMix; not vendored in the reference) ---- */
typedef struct { uint64_t seed, gamma; } orc_smgen;
uint64_t orc_mix64(uint64_t z);
uint64_t orc_mix_gamma(uint64_t z);
void     orc_mk_smgen(uint64_t s, orc_smgen* g);            /* mkSMGen */
uint64_t orc_next_word64(orc_smgen* g);                     /* nextWord64 */
double   orc_next_double(orc_smgen* g);                     /* nextDouble = (w >> 11) * 2^-53 */
/* splitmix-distributions-0.9: bernoulli p = (nextDouble < p); stdNormal via inverse normal CDF
 * (Data.Number.Erf.invnormcdf: Acklam + one Halley step).  UNVERIFIED restatement (from memory). */
double   orc_invnormcdf(double p);
double   orc_std_normal(orc_smgen* g);

/* ---- vector algebra: src/Data/RPTree/Internal.hs ---- */
double orc_inner_sd(int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d); /* :369-382 */
void   orc_project_all(int64_t nz, const int32_t* idx, const double* val, const double* X, int64_t n, int64_t d, double* out);
double orc_inner_ss(int64_t nz1, const int32_t* i1, const double* v1,
                    int64_t nz2, const int32_t* i2, const double* v2);                            /* :351-366 */
double orc_inner_dd(const double* u, const double* v, int64_t d);                                  /* :384-385 */
double orc_metric_dd_l2(const double* u, const double* v, int64_t d);                              /* :403-406 */
double orc_metric_ss_l2(int64_t nz1, const int32_t* i1, const double* v1,
                        int64_t nz2, const int32_t* i2, const double* v2);                        /* :389-393 */
double orc_metric_sd_l2(int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d); /* :396-400 */
/* binSDD (+)/(-) with the "stop when either operand is exhausted" quirk, :455-470.  Returns length. */
int64_t orc_sum_sd (int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d, double* out);
int64_t orc_diff_sd(int64_t nz, const int32_t* idx, const double* val, const double* x, int64_t d, double* out);

/* rpTreeCfg, src/Data/RPTree/Conduit.hs:132-141 */
void orc_rptree_cfg(int64_t minl, int64_t n, int64_t d, int64_t* maxd, int64_t* nchunk, double* pnz);

/* ---- hyperplanes: Gen.hs:148-195 under Batch.hs:57-63 / Conduit.hs:114-121 draw order ---- */
/* CSR over (tree, level): off has T*maxd+1 entries.  Two-phase: call with idx=val=NULL to size. */
int64_t orc_gen_hyperplanes(uint64_t seed, int32_t T, int32_t maxd, double pnz, int32_t dim,
                            int64_t* off, int32_t* idx, double* val);

/* ---- forest ---- */
typedef struct orc_forest orc_forest;
/* forestBatch / createMulti (Internal.hs:223-240): one chunk.  X is n x d row-major. */
orc_forest* orc_forest_new(const double* X, int64_t n, int32_t d, int32_t T, int32_t maxd, int32_t minl,
                           const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val);
/* forest with chunking (Conduit.hs:157-176, Internal.hs:257-297 incl. the Bin case). chunk >= n == batch */
orc_forest* orc_forest_new_chunked(const double* X, int64_t n, int32_t d, int32_t T, int32_t maxd, int32_t minl,
                                   int64_t chunk,
                                   const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val);
/* same over SVector data points (CSR, ascending indices; pointers are borrowed): projections by innerSS
 * (Internal.hs:351-366), distances by metricSDL2 (dense query) / metricSSL2 (sparse query). chunk < 1 == batch */
orc_forest* orc_forest_new_sparse(int64_t n, int32_t d, const int64_t* sp_off, const int32_t* sp_idx, const double* sp_val,
                                  int32_t T, int32_t maxd, int32_t minl, int64_t chunk,
                                  const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val);
/* orc_forest_new with the (independent) trees built on nthreads host threads; identical result. */
orc_forest* orc_forest_new_mt(const double* X, int64_t n, int32_t d, int32_t T, int32_t maxd, int32_t minl,
                              const int64_t* hp_off, const int32_t* hp_idx, const double* hp_val, int32_t nthreads);
void orc_forest_free(orc_forest* f);

/* Canonical flat export of tree t: nodes in BFS (level-major, left-to-right) order.
 * For node g: child[g] = BFS id of left child (right = +1) or -1 for a Tip; thr/mlo/mhi valid for Bin;
 * seg_start/seg_size = the slice of perm (leaves concatenated left to right) under node g.
 * depth[g] = level.  Returns the node count; call with NULL arrays to size. */
int64_t orc_tree_export(const orc_forest* f, int32_t t, int64_t* child, int32_t* depth, double* thr, double* mlo, double* mhi,
                        int64_t* seg_start, int64_t* seg_size, uint32_t* perm);
int64_t orc_tree_size(const orc_forest* f, int32_t t);  /* treeSize, RPTree.hs:362-363 */

/* candidates (RPTree.hs:297-314): ids in result order.  Returns count (call with ids=NULL to size). */
int64_t orc_candidates(const orc_forest* f, int32_t t, const double* q, uint32_t* ids, int64_t cap);
/* knn (RPTree.hs:174-176) dedup=0; knnPQ-like (RPTree.hs:187-194,224-227: dedup by distance equality,
 * first in candidate order kept) dedup=1.  Returns the number of results (<= k). */
int64_t orc_knn(const orc_forest* f, const double* q, int32_t k, int32_t dedup, double* dist, uint32_t* ids);
/* the same three with an SVector query (nz, idx, val) */
int64_t orc_candidates_sq(const orc_forest* f, int32_t t, int64_t qnz, const int32_t* qidx, const double* qval, uint32_t* ids, int64_t cap);
int64_t orc_knn_sq(const orc_forest* f, int64_t qnz, const int32_t* qidx, const double* qval, int32_t k, int32_t dedup, double* dist, uint32_t* ids);
double  orc_recall_sq(const orc_forest* f, int64_t qnz, const int32_t* qidx, const double* qval, int32_t k);
/* knnH (RPTree.hs:199-217) over candidatesH (RPTree.hs:318-341).  Returns the number of results (may exceed k: the
 * first popped leaf is always taken whole); at most cap are written.  Ties between equal priorities: unpinned in the
 * reference (heaps internals); here (tree, leaf position). */
int64_t orc_knn_h(const orc_forest* f, const double* q, int32_t k, double* dist, uint32_t* ids, int64_t cap);
int64_t orc_knn_h_sq(const orc_forest* f, int64_t qnz, const int32_t* qidx, const double* qval, int32_t k, double* dist, uint32_t* ids, int64_t cap);
/* recallWith (RPTree.hs:265-282): mean over trees of |cands(t) n topk| / k; point identity = row id. */
double  orc_recall(const orc_forest* f, const double* q, int32_t k);
/* same value, brute-force distances evaluated once instead of once per tree (falls back to orc_recall's path when a
 * distance tie straddles rank k or a tree does not hold every row) -- makes full-size recall checks affordable */
double  orc_recall_shared(const orc_forest* f, const double* q, int32_t k);
/* exact brute-force k nearest (stable by row id) -- used for forest-level recall */
void    orc_brute_knn(const double* X, int64_t n, int32_t d, const double* q, int32_t k, double* dist, uint32_t* ids);
/* use libm pow(x,2.0) instead of x*x for the squared terms (GHC `** 2`, Internal.hs:404). default 0 */
void    orc_set_use_pow(int on);

#ifdef __cplusplus
}
#endif
#endif
