/*
 * rpforest.h -- C ABI of the B200-native random-projection-forest engine (librpforest.so).
 *
 * This is the drop-in boundary for the hot path of ocramz/rp-tree (Haskell, package rp-tree-0.7.1):
 *   forestBatch / forest  ->  candidates / knn / knnPQ  ->  recallWith.
 * The reference has no FFI of its own (pure Haskell); the boundary is the export list of Data.RPTree
 * (src/Data/RPTree.hs:50-113).  A Haskell shim (hs/Data/RPTree/CUDA.hs) keeps those
 * signatures and marshals to the entry points below with `foreign import ccall safe`.
 *
 * Conventions
 *  - plain C: pointers + sizes only.  All pointers are HOST memory owned by the caller unless a
 *    parameter is named *_dev.  Outputs are caller-allocated.
 *  - every call returns RPF_OK (0) or a negative rpf_status; rpf_last_error(h) gives the message.
 *  - a handle is bound to one CUDA device and is not re-entrant (one host thread at a time).
 *  - there is NO CPU fallback: if no CUDA device is usable, rpf_create fails.
 *  - point identity is the uint32 row number in X (the shim maps rows back to `Embed` payloads).
 *  - all arithmetic is IEEE binary64 with separate mul/add roundings (no FMA), in the reference's
 *    evaluation order, so thresholds, margins, leaf sets and knn id lists are bit-exact w.r.t.
 *    the reference algorithm given the same hyperplanes.
 */
#ifndef RPFOREST_H
#define RPFOREST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rpf_handle rpf_handle;

typedef enum {
    RPF_OK = 0,
    RPF_ERR_CUDA = -1,          /* CUDA runtime error (message has the cudaError string) */
    RPF_ERR_ARG = -2,           /* invalid argument */
    RPF_ERR_STATE = -3,         /* call order violated (e.g. build before set_points) */
    RPF_ERR_UNSUPPORTED = -4,   /* valid in the reference, not yet implemented here */
    RPF_ERR_NOMEM = -5
} rpf_status;

/* ---- lifecycle ---------------------------------------------------------------------------------- */
/* Creates an engine bound to CUDA device `device` (owns a stream and all device buffers).  Replaces nothing in the
 * reference (pure values, GC-owned); the handle is what the shim's RPForest value wraps. */
int  rpf_create(rpf_handle** out, int device);
void rpf_destroy(rpf_handle* h);
const char* rpf_last_error(const rpf_handle* h);
/* ABI version of this header. */
int  rpf_abi_version(void);

/* ---- multi-GPU: trees sharded in contiguous blocks over W GPUs, data replicated ---------------------
 * createMulti maps over the IntMap of trees (Internal.hs:234-240), knn folds per-tree candidates in ascending tree order
 * (RPTree.hs:174-176), recallWith is a mean over trees (RPTree.hs:265-268): trees are independent, so GPU r builds and
 * queries trees [t0_r, t0_r + T_r) and the results equal the single-GPU results bit for bit.  The handle owns one NCCL
 * communicator per GPU (libnccl.so.2 is loaded on first use); all exchanges run inside the engine on its own streams:
 *   points  -- every GPU uploads 1/W of each row block over its own PCIe link, NVLink all-gather, projection overlapped;
 *   knn     -- per-GPU top-k lists -> ONE packed all-gather -> merge kernel ((distance, GPU, position) = tree order);
 *   recall  -- brute-force truth sharded by query, all-gather, per-GPU hit sums added in GPU (= tree) order.
 *
 * (1) ONE process, n GPUs (what a Haskell host uses): rpf_create_multi.  The returned handle is accepted by every entry
 *     point of this header with the single-GPU meaning: rpf_set_hyperplanes takes the WHOLE forest and shards it,
 *     rpf_tree_export(t) takes the forest-wide tree index, rpf_forest_export fills [T][..] arrays in tree order, rpf_knn /
 *     rpf_recall return forest-wide results (recall_sum = sum over ALL trees).  One host thread per GPU issues the work.
 *     Not available on such a handle: rpf_set_points_device, rpf_knn_h, rpf_knn_dev / rpf_merge_topk(_dev);
 *     rpf_forest_save / _load write / read one file per GPU (path + ".gpu<r>of<n>").  n_gpus == 1 == rpf_create.
 * (2) one process per GPU (SPMD, e.g. torchrun): every process creates its own handle with rpf_create, then joins with
 *     rpf_comm_init_rank BEFORE setting points (id: 128 bytes from rpf_comm_unique_id on one rank, distributed by the
 *     caller).  Each rank passes only ITS trees to rpf_set_hyperplanes (contiguous blocks in rank order) and ALL ranks
 *     call rpf_set_points / rpf_build_from_host / rpf_knn / rpf_recall collectively with the same arguments; every rank
 *     receives the forest-wide knn / recall result.  rpf_set_points reads only rows [r*per, (r+1)*per), per = ceil(n/W),
 *     of X on rank r (the other rows need not be valid memory); rpf_build_from_host reads 1/W of every row block. */
int rpf_create_multi(rpf_handle** out, const int* gpu_ids, int n_gpus);
int rpf_comm_unique_id(void* id128);
int rpf_comm_init_rank(rpf_handle* h, int32_t world, int32_t rank, const void* id128);
/* GPUs behind this handle: n of rpf_create_multi, world of rpf_comm_init_rank, else 1. */
int rpf_num_gpus(const rpf_handle* h);

/* ---- data: V.Vector (Embed DVector Double x)  (src/Data/RPTree/Internal.hs:56-63,122-126) -------- */
/* X: n x d row-major doubles (one DVector per row).  Copied to the device; caller keeps ownership. */
int rpf_set_points(rpf_handle* h, const double* X, int64_t n, int32_t d);
/* Same, but X_dev is a device pointer on this handle's device; it is borrowed (must outlive the handle's use). */
int rpf_set_points_device(rpf_handle* h, const double* X_dev, int64_t n, int32_t d);

/* ---- data: V.Vector (Embed SVector Double x)  (Internal.hs:92-97; the reference bench's own data type,
 *      bench/time/Main.hs:77,113-122) -------------------------------------------------------------------------- */
/* n SVectors of dimension d as CSR rows (off[n+1]; idx strictly ascending per row).  The engine keeps a dense n x d
 * image plus every row's last stored component.  Projections are innerSS (Internal.hs:351-366) -- identical to innerSD
 * on the dense image for finite data; distances are metricSDL2 / metricSSL2 (Internal.hs:389-400) INCLUDING the
 * reference's quirk: binSDD / binSS stop when either operand is exhausted (Internal.hs:432-470), so components past the
 * sparse operand's last stored index are ignored. */
int rpf_set_points_sparse(rpf_handle* h, int64_t n, int32_t d, const int64_t* off, const int32_t* idx, const double* val);
int rpf_points_are_sparse(const rpf_handle* h);
/* Host-only helper for SVector QUERIES: CSR rows -> dense nq x d image Q + q_last[i] = last stored component (-1: none). */
int rpf_densify_rows(int64_t nq, int32_t d, const int64_t* off, const int32_t* idx, const double* val, double* Q, int32_t* q_last);

/* ---- hyperplanes: one SVector per (tree, level)  (Internal.hs:92-93,172-175) ---------------------- */
/* PRIMARY path: the Haskell host draws rvss with the real `sample seed (replicateM ntrees (V.replicateM
 * maxd (sparse pnz dim stdNormal)))` (src/Data/RPTree/Batch.hs:57-63, Conduit.hs:114-121) and passes them
 * here as CSR over (tree-major, level-minor): off has T*maxDepth+1 entries, idx increasing within a row.
 * For multi-GPU tree sharding pass only this rank's trees. */
int rpf_set_hyperplanes(rpf_handle* h, int32_t T, int32_t maxDepth,
                        const int64_t* off, const int32_t* idx, const double* val);
/* Convenience: regenerate the hyperplanes from the seed in C (SplitMix64 core verified against the
 * splitmix haddock vectors; the normal sampler of splitmix-distributions-0.9 is restated from memory
 * and UNVERIFIED -- see DESIGN.md).  Replaces Gen.hs:148-195 under Batch.hs:59-61.
 * Draws the full forest of T_total trees and keeps trees [t_first, t_first + T_local). */
int rpf_gen_hyperplanes(rpf_handle* h, uint64_t seed, int32_t T_total, int32_t maxDepth, double pnz, int32_t d,
                        int32_t t_first, int32_t T_local);
int64_t rpf_hyperplane_nnz(const rpf_handle* h);
/* Host-only sampler (no handle, no GPU needed): same draw as rpf_gen_hyperplanes for all T trees.
 * Returns nnz; call with idx = val = NULL to size (off, if given, has T*maxDepth+1 entries). */
int64_t rpf_sample_hyperplanes(uint64_t seed, int32_t T, int32_t maxDepth, double pnz, int32_t d,
                               int64_t* off, int32_t* idx, double* val);
/* rpTreeCfg (src/Data/RPTree/Conduit.hs:132-141): default maxDepth, chunk size and pnz. Host-only. */
void rpf_rptree_cfg(int64_t minLeaf, int64_t n, int64_t d, int64_t* maxDepth, int64_t* chunk, double* pnz);
int rpf_get_hyperplanes(const rpf_handle* h, int64_t* off, int32_t* idx, double* val);

/* ---- build: forestBatch / treeBatch (Batch.hs:29-63) == createMulti/create/insert Tip-case
 *      (Internal.hs:217-240,287-297) with partitionAtMedian (Internal.hs:484-505) ---------------------- */
int rpf_build(rpf_handle* h, int32_t maxDepth, int32_t minLeaf);
/* forestBatch with the points still in HOST memory: rpf_set_points + rpf_build in one call; the upload runs in row blocks
 * on a second stream and the projection kernel starts on the rows that have arrived (pass page-locked X for full PCIe
 * speed).  Same result as rpf_set_points followed by rpf_build. */
int rpf_build_from_host(rpf_handle* h, const double* X, int64_t n, int32_t d, int32_t maxDepth, int32_t minLeaf);
/* forest / tree (Conduit.hs:58-121; insertMulti/insert, Internal.hs:243-297): the rows of X arrive in chunks of
 * `chunk` points, in row order.  chunk >= n is identical to rpf_build (one insert into an empty Tip).  chunk < n runs
 * the reference's streaming update per chunk: every Bin a chunk passes through gets thr' = (thr0 + thr)/2 and
 * margin' = margin0 <> margin with the CHUNK's positional median (Internal.hs:274-285); every Tip gets xs <> xs0 and is
 * split again once it outgrows minLeaf (Internal.hs:287-297).  The reference's quirk is kept: an empty piece reaching a
 * Bin replaces that subtree by an empty Tip (Internal.hs:279), dropping its points -- see rpf_points_lost.
 * Limits: a Tip that must be re-split may hold at most 8192 points (minLeaf <= 4095 in practice). */
int rpf_build_chunked(rpf_handle* h, int32_t maxDepth, int32_t minLeaf, int64_t chunk);
/* The same fold, driven by the caller AS THE CHUNKS ARRIVE (Conduit.hs:157-176: `chunksOf n .| foldl insertMulti im0`;
 * one rpf_insert_chunk == one insertMulti, Internal.hs:243-255).  Neither n nor the number of chunks is known in advance
 * and the chunks may have any sizes (the reference's chunksOf gives equal chunks and a shorter last one).
 *   rpf_insert_begin : every tree = `Tip () mempty` (Conduit.hs:166-168); needs the hyperplanes (rpf_set_hyperplanes /
 *                      rpf_gen_hyperplanes); replaces the handle's points and forest.
 *   rpf_insert_chunk : X_chunk = m x d row-major HOST rows; they get the row ids n_so_far .. n_so_far + m - 1.  After every
 *                      call the handle holds a complete forest over all rows inserted so far -- every query / export /
 *                      checkpoint entry point works between chunks -- identical to rpf_build_chunked over the same rows
 *                      with the same chunk boundaries.  m = 0 is a no-op chunk.
 *   rpf_insert_end   : closes the session (frees the per-level key store); points and forest stay, as after a build.
 * Any other call that replaces the points or the hyperplanes closes the session and discards its points; rpf_build /
 * rpf_build_chunked / rpf_build_from_host are refused (RPF_ERR_STATE) while a session is open.
 * Same limit as rpf_build_chunked (a Tip that re-splits holds <= 8192 points); on that error the session is closed. */
int rpf_insert_begin(rpf_handle* h, int32_t d, int32_t maxDepth, int32_t minLeaf);
int rpf_insert_chunk(rpf_handle* h, const double* X_chunk, int64_t m);
int rpf_insert_end(rpf_handle* h);

/* ---- result structure: RPT Bin/Tip (Internal.hs:139-148) as flat arrays ---------------------------- */
/* The topology (which nodes exist, their sizes) is a pure function of (n, minLeaf, maxDepth) because the
 * split is positional (first n div 2 of the stable-sorted node go left, Internal.hs:495,503); it is the
 * same for every tree.  Nodes are numbered in BFS order (level-major, left to right). */
int64_t rpf_num_nodes(const rpf_handle* h);
int32_t rpf_num_trees(const rpf_handle* h);
int32_t rpf_hyperplane_depth(const rpf_handle* h);     /* hyperplanes stored per tree (>= maxDepth of the build) */
int rpf_points_shape(const rpf_handle* h, int64_t* n, int32_t* d);
/* child[g] = BFS id of the left child (right = +1) or -1 for a Tip; seg_start/seg_size = slice of perm
 * holding the points under node g.  Any pointer may be NULL. */
int rpf_topology(const rpf_handle* h, int64_t* child, int32_t* depth, int64_t* seg_start, int64_t* seg_size);
/* Host-only: the topology for (n, maxDepth, minLeaf) without a handle.  Returns the node count; arrays may be NULL. */
int64_t rpf_topology_plan(int64_t n, int32_t maxDepth, int32_t minLeaf, int64_t* child, int32_t* depth,
                          int64_t* seg_start, int64_t* seg_size);
/* Host-only: the tree shape after a streaming build with the given chunk size (chunk >= n: same as rpf_topology_plan).
 * points_lost (may be NULL) receives the number of points the reference's empty-piece rule drops. */
int64_t rpf_topology_plan_chunked(int64_t n, int32_t maxDepth, int32_t minLeaf, int64_t chunk, int64_t* child, int32_t* depth,
                                  int64_t* seg_start, int64_t* seg_size, int64_t* points_lost);
/* Points dropped by the last rpf_build_chunked (0 for rpf_build).  The slots of perm past seg_size[0] hold 0xffffffff. */
int64_t rpf_points_lost(const rpf_handle* h);
/* 1 if every leaf's internal order equals the reference's (always, unless a leaf is larger than the
 * shared-memory capacity set by rpf_set_bottom_cap; leaf SETS are exact regardless). */
int rpf_leaf_order_exact(const rpf_handle* h);
/* Per tree: thr/mlo/mhi[num_nodes] (_rpThreshold, Margin low/high; valid where child>=0) and
 * perm[n] = row ids, leaves concatenated left to right, each leaf in the reference's order. */
int rpf_tree_export(rpf_handle* h, int32_t t, double* thr, double* mlo, double* mhi, uint32_t* perm);
/* Whole forest in one call: thr/mlo/mhi[T][num_nodes], perm[T][n] (tree-major; any pointer may be NULL).
 * The copies run back to back on the engine's stream; pass page-locked buffers for full PCIe speed. */
int rpf_forest_export(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm);
/* Export sink: host buffers (same shapes as rpf_forest_export's, ideally page-locked; perm must be non-NULL, NULLs clear
 * the sink) that rpf_build_from_host fills WHILE it builds -- the bottom phase runs in tree groups and every group's
 * slice of perm is downloaded on a second stream as soon as it is final -- so the RPForest value forestBatch returns
 * (Batch.hs:48-63) is on the host ~2 ms after the last kernel instead of after a separate 153 MB download.
 * rpf_forest_export with exactly these pointers then only waits for the stream.  The buffers must stay valid until the
 * sink is cleared or the handle destroyed; their content is undefined between a build and the matching export. */
int rpf_set_export_sink(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm);

/* ---- checkpoint (the engine-side counterpart of serialiseRPForest / deserialiseRPForest, Internal.hs:185-196) -------- */
/* One flat little-endian file: hyperplanes, topology, thr/mlo/mhi, perm and -- with_points != 0 -- the data points
 * (like the reference's serialised forest, which carries the Embed values in its leaves).  rpf_forest_load restores a
 * queryable forest without rebuilding; a checkpoint without points needs the same points set on the handle first. */
int rpf_forest_save(rpf_handle* h, const char* path, int32_t with_points);
int rpf_forest_load(rpf_handle* h, const char* path);

/* ---- queries --------------------------------------------------------------------------------------- */
/* candidates (src/Data/RPTree.hs:293-314) for tree t (t >= 0) or for all trees concatenated tree-major
 * (t == -1, i.e. `fold ((`candidates` q) <$> tts)`, RPTree.hs:176).  Q: nq x d row-major.
 * Two calls: counts -> caller sizes ids -> fill.  off_out has nq+1 entries (CSR). */
int rpf_candidates_count(rpf_handle* h, const double* Q, int64_t nq, int32_t t, int64_t* off_out);
int rpf_candidates(rpf_handle* h, const double* Q, int64_t nq, int32_t t, const int64_t* off, uint32_t* ids);
/* knn metricL2 k (RPTree.hs:168-176; dedup=0: duplicates across trees kept, exactly as the reference)
 * knnPQ metricL2 k (RPTree.hs:181-194,224-227; dedup=1: one result per distinct distance).
 * dist/ids: nq x k row-major; count[q] = number of valid results (<= k).  Distances are
 * sqrt(sum_j (x_j - q_j)^2), left-fold sum, correctly rounded squares (Internal.hs:403-406). */
int rpf_knn(rpf_handle* h, const double* Q, int64_t nq, int32_t k, int32_t dedup,
            double* dist, uint32_t* ids, int32_t* count);
/* knnH metricL2 k (RPTree.hs:199-217, candidatesH :318-341): the leaves the descent reaches are ranked by their margin
 * priority (smallest margin distance met on the way down); leaves are taken in increasing priority and PREPENDED to the
 * result while the running total stays <= k (the first non-empty one is always taken).  As in the reference the result
 * is neither sorted by distance nor cut to k: count[q] <= cap entries per query, cap >= rpf_knn_h_capacity(h, k) =
 * max(k, largest leaf).  q_last: NULL, or as in rpf_knn_s.  Order among leaves of EQUAL priority is an internal of the
 * `heaps` package in the reference (unpinned); here: tree index, then leaf position.  Limit: trees x (most leaves one
 * query reaches in one tree) <= 9216. */
int64_t rpf_knn_h_capacity(const rpf_handle* h, int32_t k);
int rpf_knn_h(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int64_t cap,
              double* dist, uint32_t* ids, int32_t* count);
/* recallWith metricL2 forest k q (RPTree.hs:259-282): mean over this handle's trees of
 * |candidates(t,q) /\ true-top-k| / k.  recall_sum[q] = SUM over local trees (divide by the global
 * tree count after reducing across GPUs).
 * Deviation after a LOSSY streaming build (rpf_points_lost > 0): the truth here ranks all n rows, ties by row id; the
 * reference's recallWith1 ranks `points tt`, i.e. only the points that tree still holds, in leaf order with a stable sort
 * (RPTree.hs:265-282) -- dropped rows can enter the truth set here.  With no points lost and no exact distance tie at rank k
 * the two agree (tested against the oracle). */
int rpf_recall(rpf_handle* h, const double* Q, int64_t nq, int32_t k, double* recall_sum);   /* multi-GPU handle / communicator rank: sum over ALL trees */
/* Exact brute-force k nearest rows (ties by row id): ground truth for forest-level recall. */
int rpf_brute_knn(rpf_handle* h, const double* Q, int64_t nq, int32_t k, double* dist, uint32_t* ids);

/* The same three for SVector data with the query's representation made explicit: q_last == NULL means DVector queries
 * (metricSDL2; identical to rpf_knn / rpf_recall / rpf_brute_knn), otherwise SVector queries given as their dense image
 * plus q_last (rpf_densify_rows) -> metricSSL2.  SVector queries against DVector data are rejected (no such Inner
 * instance in the reference, Internal.hs:322-341). */
int rpf_knn_s(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int32_t dedup,
              double* dist, uint32_t* ids, int32_t* count);
int rpf_recall_s(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* recall_sum);
int rpf_brute_knn_s(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* dist, uint32_t* ids);

/* ---- multi-GPU: merge per-GPU top-k lists (trees sharded in contiguous blocks, rank-major) ---------- */
/* dist/ids: G x nq x k, count: G x nq (rank-major).  Result ordered by (distance, rank, position), which
 * equals the single-GPU order of rpf_knn over the whole forest. */
int rpf_merge_topk(rpf_handle* h, int32_t G, int64_t nq, int32_t k, int32_t dedup,
                   const double* dist, const uint32_t* ids, const int32_t* count,
                   double* dist_out, uint32_t* ids_out, int32_t* count_out);

/* Device-resident form of the exchange: rpf_knn_dev leaves this rank's lists in caller-provided DEVICE buffers (nq x k
 * doubles, nq x k uint32, nq int32 on this handle's device; complete when the call returns), the caller all-gathers them
 * over NCCL / NVLink, and rpf_merge_topk_dev merges the gathered rank-major DEVICE lists; only the merged result
 * crosses PCIe.  q_last as in rpf_knn_s (NULL for DVector queries). */
int rpf_knn_dev(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int32_t dedup,
                double* dist_dev, uint32_t* ids_dev, int32_t* count_dev);
int rpf_merge_topk_dev(rpf_handle* h, int32_t G, int64_t nq, int32_t k, int32_t dedup,
                       const double* dist_dev, const uint32_t* ids_dev, const int32_t* count_dev,
                       double* dist_out, uint32_t* ids_out, int32_t* count_out);

/* ---- measurement hooks (CUDA events on the engine's own stream) ------------------------------------ */
/* Device time in ms of the most recent rpf_build / rpf_knn / rpf_recall kernels (events on the stream
 * the kernels were launched on; excludes host<->device copies of the call's arguments). */
double rpf_last_device_ms(const rpf_handle* h);
/* Per-phase profile of the last call when profiling is enabled (adds event records between phases).
 * Phases: see rpf_phase_name(i).  Returns the number of phases; ms/launches may be NULL. */
int rpf_set_profiling(rpf_handle* h, int on);
int rpf_get_profile(const rpf_handle* h, double* ms, int64_t* launches, int cap);
const char* rpf_phase_name(int i);
/* Total kernel launches issued by this handle since creation. */
int64_t rpf_launch_count(const rpf_handle* h);
/* Named options (tuning knobs and test hooks; results are identical under every setting):
 *   "lean_top" (0/1, default 1: 0 = generic top-phase compact / relabel kernels only), "fuse_relabel_hist" (0/1, default 1: top-phase
 *   relabel of level l fused with the histogram of level l + 1), "hist_big_chunk" (0/1, default 1), "fused_top" (bit mask, default 1),
 *   "fused_pick_min_tg" (default 16), "top_chunk_hist" / "top_chunk_compact" / "top_chunk_relabel" (points per CTA, 0 = automatic),
 *   "branches" (concurrent tree blocks per build, 0 = automatic), "cuda_graph" (0/1), "project_variant", "project_prefetch",
 *   "project_pipe_maxh", "force_generic_bottom" (0/1), "bottom_words64" (0/1), "bottom_select" (0/1, default 0), "knn_filter32" (0/1,
 *   default 1), "knn_f32_stages" / "knn_f32_rows" / "knn_f32_buf" / "knn_f32_sreg", "force_simple_knn", "force_simple_topk",
 *   "no_query_order", "rerank_gemm" (0/1/2), "release_workspace" (free the cached device workspace now).  Unknown names: RPF_ERR_ARG. */
int rpf_set_option(rpf_handle* h, const char* name, int64_t value);
/* Tuning knob: bottom-phase shared-memory capacity in points (256, 1024, 4096 or 8192). */
int rpf_set_bottom_cap(rpf_handle* h, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif
