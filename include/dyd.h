/* dyd.h -- C ABI of libdyd.so, the B200 (sm_100a) hot path of Deal-Yolo-Daya's
 * data-processing pipeline.
 *
 * The reference (/root/reference/src/deal_yolo_data/core/processor.py) has no FFI:
 * its boundary is the Python step-function surface imported by the Streamlit page
 * (ui/pages/processing.py:25-38).  The Python drop-in (deal_yolo_daya_b200/
 * processor.py) keeps those signatures and calls the entry points below through
 * ctypes; each entry point states which reference lines it replaces.
 *
 * Conventions
 *   - plain pointers + int64 counts; no torch / C++ types in any signature;
 *   - "d_" pointers are DEVICE pointers on the current CUDA device, "h_" pointers
 *     are HOST pointers (pinned gives full PCIe speed, pageable is accepted);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     device entry points are stream-ordered and never synchronise the host;
 *   - the library owns no memory: scratch comes from the caller as a workspace
 *     sized by the matching dyd_*_workspace_bytes() query;
 *   - return value: 0 ok; <0 invalid argument (DYD_E_*); >0 a cudaError_t.
 *     dyd_last_error() gives the thread-local message of the last failure;
 *   - re-entrant and callable from any host thread.  The only library-owned state is the per-device
 *     CUDA memory pool behind the host-buffer entry points (created on first use under a mutex,
 *     returned to the driver by dyd_host_release()); everything else lives in caller memory.
 *
 * Data layout: a ragged CSR annotation table
 *   img_off  int64[n_img+1]   object range of each image (row)
 *   poly_off int64[n_poly+1]  vertex range of each object's polygon
 *   xy       double[2*n_vert] interleaved x,y of the VALID points only, 16-byte aligned
 *   pts      double[4*n_poly] per object (p1.x, p1.y, p2.x, p2.y) = (min_x, min_y, max_x, max_y)
 *   valid    uint8[n_poly]    0 = null bbox ({x: None, y: None} twice in the reference)
 */
#ifndef DYD_H_
#define DYD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYD_VERSION 100          /* 0.1.0 */
#define DYD_E_ARG (-1)           /* null pointer / negative count / bad enum */
#define DYD_E_ALIGN (-2)         /* pointer not aligned as required */
#define DYD_E_WORKSPACE (-3)     /* workspace too small */
#define DYD_E_IO (-4)            /* a file could not be opened / written (errno holds the reason) */

int dyd_version(void);
/* Number of CUDA kernels this library has launched in the calling process so far (statistic; bench.py reports the
 * difference over its timed region as `gpu_launches`).                                                  */
uint64_t dyd_launch_count(void);
/* Copies the calling thread's last error message (NUL terminated) into buf. */
size_t dyd_last_error(char* buf, size_t cap);

/* ---------------------------------------------------------------- K1 ------
 * polygon -> two corner points.  Replaces get_bbox_points, processor.py:252-260:
 * four independent builtin min()/max() scans over the valid points, i.e. left
 * folds that replace the running value only on a strict comparison (first of
 * equal values -- and of 0.0 / -0.0 -- wins; a NaN survives only from position 0).
 * Outputs: pts (above), valid, and optionally arg int32[4*n_poly] = the vertex
 * index inside the polygon that supplied each of the four values (lets the host
 * re-emit the original JSON number, `10` vs `10.0`).  d_arg may be NULL.        */
int dyd_bbox_minmax(const int64_t* d_poly_off, const double* d_xy, int64_t n_poly,
                    double* d_pts, uint8_t* d_valid, int32_t* d_arg, void* stream);

/* ---------------------------------------------------------------- K2 ------
 * box-count + any-pair IoU quality flag per image.  Replaces extract_boxes /
 * calculate_iou / meet_conditions, processor.py:328-376:
 *   boxes = objects of the image up to (not including) the first one whose
 *           valid flag is 0 (the TypeError a null bbox raises ends the scan);
 *           each box = (min(p1x,p2x), min(p1y,p2y), max(p1x,p2x), max(p1y,p2y));
 *   high  = len(boxes) >= min_boxes and any i<j: iou(boxes[i], boxes[j]) >= thr
 *   iou   = fp64, separately rounded: inter = max(0,dx)*max(0,dy); 0.0 if inter==0;
 *           union = (area1 + area2) - inter; inter/union, or 0.0 if union == 0.
 * d_valid may be NULL (all valid).  d_count receives len(boxes).
 * Workspace: dyd_iou_workspace_bytes(n_img) (worklist of crowded images).       */
size_t dyd_iou_workspace_bytes(int64_t n_img);
int dyd_iou_filter(const int64_t* d_img_off, const double* d_pts, const uint8_t* d_valid,
                   int64_t n_img, int64_t min_boxes, double thr,
                   uint8_t* d_high, int32_t* d_count,
                   void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------- K1+K2 ------
 * Fused polygon->bbox + IoU flag (processor.py:252-260 then :328-376 without the
 * CSV/JSON round trip between the two steps): one pass over the vertices, boxes
 * written once and not re-read.  Same outputs as the two calls above.           */
int dyd_bbox_iou_fused(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                       int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                       double* d_pts, uint8_t* d_valid, int32_t* d_arg,
                       uint8_t* d_high, int32_t* d_count,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/* Same call restricted to at most `max_ctas` persistent CTAs (one per SM; <= 0 or >= 148: all SMs).  The
 * kernel claims its image segments dynamically, so a caller that runs another stream next to it -- the URL
 * hash / dedup / exchange chain of the pipeline, which depends on the `source` column only -- can leave that
 * stream a few SMs and have both finish together (bench.py).                           */
int dyd_bbox_iou_fused_ex(const int64_t* d_img_off, const int64_t* d_poly_off, const double* d_xy,
                          int64_t n_img, int64_t n_poly, int64_t min_boxes, double thr,
                          double* d_pts, uint8_t* d_valid, int32_t* d_arg,
                          uint8_t* d_high, int32_t* d_count,
                          void* d_workspace, size_t workspace_bytes, int32_t max_ctas, void* prepass_done_event, void* stream);
/* prepass_done_event (cudaEvent_t as void*, may be NULL) is recorded on `stream` between the descriptor pre-pass and the
 * persistent kernel: a second stream that waits for it starts when the persistent CTAs are being placed, so -- with
 * `stream` created at a higher priority -- they get their SMs first and the other stream's kernels fill what is left.
 * dyd_fused_cta_times: diagnostics, host buffer uint64[n <= 296] = (start, end) globaltimer of the CTAs of the last
 * staged launch in this process.                                                                          */
int dyd_fused_cta_times(uint64_t* h_times, int32_t n);
/* Diagnostics: tiles of each staging mode (0 bulk-copy staged, 1 direct loads, 2 deferred crowded image) that the
 * last fused call on this workspace produced; d_counts3 uint64[3].                                   */
int dyd_fused_tile_modes(const void* d_workspace, int64_t n_img, uint64_t* d_counts3, void* stream);

/* ---------------------------------------------------------------- K0 ------
 * 64-bit hash of each string of an Arrow-style (offsets, bytes) column; the
 * `source` URLs of processor.py:140 / :194.  Definition in DESIGN.md §4.K0.     */
int dyd_hash_strings(const int64_t* d_off, const uint8_t* d_bytes, int64_t n,
                     uint64_t* d_hash, void* stream);

/* ---------------------------------------------------------------- K4 ------
 * Row keep-mask of df.drop_duplicates(subset=["source"], keep=...), processor.py:
 * 140-144.  keep_mode 0 = "first", 1 = "last", 2 = False (drop all members of a
 * repeated group).  d_null may be NULL; all null (NaN) rows form ONE group.
 * d_rep[r] = row representing r's group (first, or last for mode 1) so the host
 * can verify every dropped row's string against it (64-bit hash collisions).     */
size_t dyd_dedup_workspace_bytes(int64_t n);
int dyd_dedup(const uint64_t* d_keys, const uint8_t* d_null, int64_t n, int keep_mode,
              uint8_t* d_keep, int64_t* d_rep,
              void* d_workspace, size_t workspace_bytes, void* stream);

/* Sharded form used after the NCCL all-to-all: rows carry explicit global ids. */
int dyd_dedup_ids(const uint64_t* d_keys, const int64_t* d_row_id, int64_t n, int keep_mode,
                  uint8_t* d_keep, int64_t* d_rep,
                  void* d_workspace, size_t workspace_bytes, void* stream);

/* Multi-GPU exchange helpers for K4 (one process per GPU; the all-to-all itself is NCCL).
 * _bucket: scatters (key, row_base + row) records into `world` buckets of `cap` records each
 *          (int64[world*cap*2], padding = id -1), owner = mix(key) mod world; *d_overflow = 1 if a
 *          bucket was too small (the caller then uses exact-size splits).  Null rows are skipped.
 * dyd_dedup_ids ignores padded records.  _pack_reply builds (id, rep | keep << 62) answers on the
 * owner; _unpack places them at rows id - row_base on the origin.                              */
int dyd_shard_bucket(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                     int64_t cap, int64_t* d_records, uint64_t* d_cursors, int32_t* d_overflow, void* stream);
int dyd_dedup_records(const int64_t* d_records, int64_t m, int keep_mode, uint8_t* d_keep, int64_t* d_rep,
                      void* d_workspace, size_t workspace_bytes, void* stream);   /* Peer-memory forms of the exchange steps (NVLink P2P, buffers mapped into every rank, e.g. by
 * torch.distributed._symmetric_memory): the scatter writes each record straight into region `my_rank` of
 * its owner's receive buffer (and notes locally, in d_sent_row[world*cap], which row went into which
 * slot); the owner writes each 8-byte answer (rep | keep << 62, -1 for padding) straight into region
 * `my_rank` of the origin's reply buffer (int64[world*cap]); the origin places the answers with
 * d_sent_row and its own cursors.  The compute kernels ARE the all-to-all.  d_peer_records / d_peer_reply:
 * device arrays of `world` pointers (own buffer included); the receive buffers must be pre-filled with
 * 0xFF padding and the ranks must synchronise before and after each step (sharding.DedupExchange).   */
int dyd_shard_bucket_p2p(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                         int32_t my_rank, int64_t cap, int64_t* const* d_peer_records, uint32_t* d_sent_row,
                         uint64_t* d_cursors, int32_t* d_overflow, void* stream);
/* mode 0: dedup answers (rep | keep << 62); mode 1: anti-join answers (kept ? 1 << 62 : first matching
 * reference row; _unpack then writes ref_row = -1 for kept rows).  reset_records != 0: the pack kernel, the
 * last reader of the received records, turns each one back into padding (id = -1) so the next step needs no
 * fill of the receive buffer.                                                                           */
int dyd_shard_pack_reply_p2p(int64_t* d_records, const uint8_t* d_keep, const int64_t* d_rep, int64_t m, int64_t cap,
                             int32_t my_rank, int64_t* const* d_peer_reply, int32_t mode, int32_t reset_records, void* stream);
int dyd_shard_unpack_p2p(const int64_t* d_reply, const uint32_t* d_sent_row, const uint64_t* d_cursors, int32_t world,
                         int64_t cap, int64_t n, uint8_t* d_keep, int64_t* d_rep, int32_t mode, void* stream);
/* Joint exchange (dedup + anti-join answers of the same main records, sharding.UrlFilterExchange): both 8-byte answers
 * travel as ONE 16-byte store into region `my_rank` of the origin's reply buffer (int64[2 * world * cap], 16-byte
 * aligned); _unpack2 places them into the four result columns.
 * Sparse replies (sparse != 0): _bucket_p2p_defaults, the scatter of the main table, also writes the answer "kept by both"
 * (keep 1, rep = own global row, keep_ref 1, ref_row -1) for every row with coalesced stores; the origin's reply buffer holds
 * -1 everywhere (filled once; _unpack2 resets the slots it reads); the owners then store only the answers that differ
 * (a few per cent of the records cross NVLink on the way back) and _unpack2 scatters only those.                   */
int dyd_shard_bucket_p2p_defaults(const uint64_t* d_keys, const uint8_t* d_null, int64_t row_base, int64_t n, int32_t world,
                                  int32_t my_rank, int64_t cap, int64_t* const* d_peer_records, uint32_t* d_sent_row,
                                  uint64_t* d_cursors, int32_t* d_overflow, uint8_t* d_keep_dedup, int64_t* d_rep_dedup,
                                  uint8_t* d_keep_anti, int64_t* d_ref_row, void* stream);
int dyd_shard_pack_reply2_p2p(int64_t* d_records, const uint8_t* d_keep_dedup, const int64_t* d_rep_dedup,
                              const uint8_t* d_keep_anti, const int64_t* d_ref_row, int64_t m, int64_t cap, int32_t my_rank,
                              int64_t* const* d_peer_reply2, int32_t reset_records, int32_t sparse, void* stream);
int dyd_shard_unpack2_p2p(int64_t* d_reply2, const uint32_t* d_sent_row, const uint64_t* d_cursors, int32_t world,
                          int64_t cap, int64_t n, uint8_t* d_keep_dedup, int64_t* d_rep_dedup, uint8_t* d_keep_anti,
                          int64_t* d_ref_row, int32_t sparse, void* stream);
/* NCCL-transport forms: (id, answer) pairs */
int dyd_shard_pack_reply(const int64_t* d_records, const uint8_t* d_keep, const int64_t* d_rep, int64_t m,
                         int64_t* d_reply, int32_t mode, void* stream);
int dyd_shard_unpack(const int64_t* d_reply, int64_t m, int64_t row_base, int64_t n,
                     uint8_t* d_keep, int64_t* d_rep, int32_t mode, void* stream);

/* ---------------------------------------------------------------- K5 ------
 * Anti-join of processor.py:194-199: keep[r] = main value not in
 * set(ref.dropna()); null main rows are always kept.  d_ref_row[r] = first
 * reference row holding the key (-1 when kept), for collision verification.     */
size_t dyd_antijoin_workspace_bytes(int64_t n_ref);
int dyd_antijoin(const uint64_t* d_main_keys, const uint8_t* d_main_null, int64_t n_main,
                 const uint64_t* d_ref_keys, const uint8_t* d_ref_null, int64_t n_ref,
                 uint8_t* d_keep, int64_t* d_ref_row,
                 void* d_workspace, size_t workspace_bytes, void* stream);

/* Sharded form (owner side of the exchange): both tables arrive as (key, id) records in fixed-capacity buckets
 * (id < 0 = padding).  d_keep / d_ref_row are indexed by main record; d_ref_row = smallest reference id holding the
 * key (global reference row), -1 when kept; padding records get keep 0.  reset_ref != 0 turns the reference
 * records back into padding after they were read.  Workspace: dyd_antijoin_workspace_bytes(m_ref).      */
int dyd_antijoin_records(int64_t* d_ref_records, int64_t m_ref, const int64_t* d_main_records, int64_t m_main,
                         uint8_t* d_keep, int64_t* d_ref_row, void* d_workspace, size_t workspace_bytes,
                         int32_t reset_ref, void* stream);

/* ------------------------------------------------------------ K4 + K5 ------
 * Steps 2 and 3 on the same main `source` keys in one call (processor.py:140-144 and :194-199): the outputs are exactly
 * those of dyd_dedup (d_keep / d_rep) and dyd_antijoin (d_keep_ref / d_ref_row).  Large tables are scattered into common key
 * partitions and one shared-memory table per partition answers both questions; small ones run the two kernels above.
 * _records: the sharded form on (key, id) records of the exchange (dyd_dedup_records + dyd_antijoin_records).        */
size_t dyd_url_filter_workspace_bytes(int64_t n_main, int64_t n_ref);
int dyd_url_filter(const uint64_t* d_main_keys, const uint8_t* d_main_null, int64_t n_main,
                   const uint64_t* d_ref_keys, const uint8_t* d_ref_null, int64_t n_ref, int keep_mode,
                   uint8_t* d_keep, int64_t* d_rep, uint8_t* d_keep_ref, int64_t* d_ref_row,
                   void* d_workspace, size_t workspace_bytes, void* stream);
int dyd_url_filter_records(int64_t* d_ref_records, int64_t m_ref, const int64_t* d_main_records, int64_t m_main, int keep_mode,
                           uint8_t* d_keep, int64_t* d_rep, uint8_t* d_keep_ref, int64_t* d_ref_row,
                           void* d_workspace, size_t workspace_bytes, int32_t reset_ref, int64_t id_bound, void* stream);
/* id_bound: every record id (main and reference) is below this value (0 = unknown); below 2^31 the shared-memory tables keep
 * 32-bit rows (native shared atomics).                                                                      */

/* ---------------------------------------------------------------- K3 ------
 * Object-name rewrite through a lookup table, processor.py:582-602 with the
 * string rules of utils.py:659-679 folded into three per-vocabulary tables built
 * on the host: lut_new (id of the rewritten name), lut_ntok (#tokens), lut_nrep
 * (#tokens found in the mapping).  label_id < 0 = object without a name.
 * d_counters uint64[6] = total_objects, missing_name_objects, total_labels,
 * replaced_labels, replaced_objects, replaced_rows (zeroed by the call).        */
int dyd_label_lut(const int64_t* d_img_off, const int32_t* d_label_id, int64_t n_img, int64_t n_box,
                  const int32_t* d_lut_new, const int32_t* d_lut_ntok, const int32_t* d_lut_nrep,
                  int32_t n_vocab, int32_t* d_new_id, uint8_t* d_row_replaced,
                  uint64_t* d_counters, void* stream);

/* Occurrences of every vocabulary id among the objects (uint64[n_vocab], zeroed by the call);
 * with the host's token tables this gives the unmatched-label counts of processor.py:591-593. */
int dyd_label_hist(const int32_t* d_label_id, int64_t n_box, int32_t n_vocab, uint64_t* d_hist, void* stream);

/* Per-image label presence + per-box counts of summarize_yolo_label_counts (processor.py:1113-1131): d_img_hist[v] =
 * images holding at least one object with id v ("图片数量"), d_box_hist[v] = objects with id v ("标注框数量"); both
 * uint64[n_vocab], zeroed by the call; ids outside [0, n_vocab) are ignored.                              */
int dyd_label_presence(const int64_t* d_img_off, const int32_t* d_label_id, int64_t n_img, int32_t n_vocab,
                       uint64_t* d_img_hist, uint64_t* d_box_hist, void* stream);

/* ---------------------------------------------------------------- K6 ------
 * label -> category expansion of processor.py:751-775: one expanded row per
 * object whose label has a category, stably grouped by category (encounter
 * order inside a category).  Two calls: _count fills d_cat_off int64[n_cat+1]
 * (the caller reads d_cat_off[n_cat] to size the outputs) and leaves per-chunk
 * offsets in the workspace; _fill (same workspace, untouched in between) scatters
 * (image, object, category) triples.  n_cat <= 256.
 * Workspace: dyd_split_workspace_bytes(n_img, n_cat).                            */
size_t dyd_split_workspace_bytes(int64_t n_img, int32_t n_cat);
int dyd_split_count(const int64_t* d_img_off, int64_t n_img, const int32_t* d_label_id,
                    const int32_t* d_cat_of_label, int32_t n_vocab, int32_t n_cat, int64_t* d_cat_off,
                    void* d_workspace, size_t workspace_bytes, void* stream);
int dyd_split_fill(const int64_t* d_img_off, int64_t n_img, const int32_t* d_label_id,
                   const int32_t* d_cat_of_label, int32_t n_vocab, int32_t n_cat, const int64_t* d_cat_off,
                   int64_t* d_exp_img, int64_t* d_exp_box, int32_t* d_exp_cat,
                   void* d_workspace, size_t workspace_bytes, void* stream);
/* Split id (0 train / 1 val / 2 test) and shuffled position of every expanded row.
 * d_perm holds, category after category, np.random.RandomState(seed).permutation(n_c)
 * computed on the host (what DataFrame.sample(frac=1, random_state=seed) applies,
 * processor.py:800): shuffled position r takes original row perm[r]; positions
 * < n_train[c] are train, the next n_val[c] val, the rest test (:801-806).        */
int dyd_split_assign(const int64_t* d_cat_off, int32_t n_cat, const int64_t* d_perm, int64_t n_exp,
                     const int64_t* d_n_train, const int64_t* d_n_val,
                     uint8_t* d_split, int64_t* d_pos, void* stream);

/* Sharded form (rows partitioned by image over the ranks): d_cat_off / d_perm / n_exp describe the GLOBAL
 * category-grouped table; this rank owns rows d_own_lo[c] .. d_own_lo[c] + d_own_cnt[c] of category c (positions inside
 * the category; sharding.split_category_bases) and receives their split id and shuffled position at
 * d_local_off[c] + k of its own d_split / d_pos.  Every rank sweeps the whole permutation once.            */
int dyd_split_assign_range(const int64_t* d_cat_off, int32_t n_cat, const int64_t* d_perm, int64_t n_exp,
                           const int64_t* d_n_train, const int64_t* d_n_val, const int64_t* d_own_lo,
                           const int64_t* d_own_cnt, const int64_t* d_local_off, uint8_t* d_split, int64_t* d_pos, void* stream);

/* np.random.RandomState(seed).permutation(n), bit for bit (host; what DataFrame.sample(frac=1, random_state=seed) of
 * processor.py:800 / :1003 applies): MT19937 + numpy's legacy masked-rejection shuffle, with the swap targets drawn a
 * block ahead and prefetched.  h_out int64[n].  Integer seeds 0 .. 2^32-1 (what the reference passes).               */
int dyd_numpy_permutation(uint32_t seed, int64_t n, int64_t* h_out);

/* ------------------------------------------------------------ YOLO (f-3) ---
 * cx, cy, w, h of processor.py:1045-1052 for every box: ((x1+x2)/2)/W etc., fp64,
 * same operation order; ok[q] = 0 where bw <= 0 or bh <= 0 or the box is invalid. */
int dyd_yolo_normalise(const int64_t* d_img_off, const double* d_pts, const uint8_t* d_valid,
                       const double* d_img_wh, int64_t n_img, int64_t n_box,
                       double* d_cxcywh, uint8_t* d_ok, void* stream);

/* ------------------------------------------------------ host-buffer entry ---
 * The call the Python drop-in makes for a table that lives in HOST memory: the
 * image range is cut into chunks; each chunk's CSR slice is copied H2D, run
 * through the fused K1+K2 kernel and its results copied D2H, with copies and
 * kernels of consecutive chunks overlapped on internal streams.  Blocks until
 * the results are in the h_ outputs.  h_arg / h_pts may be NULL (not wanted).
 * `chunk_images` <= 0 picks a default.                                          */
int dyd_bbox_iou_host(const int64_t* h_img_off, const int64_t* h_poly_off, const double* h_xy,
                      int64_t n_img, int64_t min_boxes, double thr,
                      double* h_pts, uint8_t* h_valid, int32_t* h_arg,
                      uint8_t* h_high, int32_t* h_count, int64_t chunk_images);

/* Same call, also reporting how the staged kernel handled the chunks: h_tile_modes int64[3] = tiles that were
 * bulk-copy staged / read with direct loads / deferred to the block-per-image kernel (may be NULL).        */
int dyd_bbox_iou_host_ex(const int64_t* h_img_off, const int64_t* h_poly_off, const double* h_xy,
                         int64_t n_img, int64_t min_boxes, double thr,
                         double* h_pts, uint8_t* h_valid, int32_t* h_arg,
                         uint8_t* h_high, int32_t* h_count, int64_t chunk_images, int64_t* h_tile_modes);

/* The two host entry points cache their transient device buffers in a library-owned CUDA memory
 * pool (per device); this returns the cached blocks of the current device to the driver. */
int dyd_host_release(void);

/* Same idea for the URL column: hash + dedup (+ anti-join when n_ref > 0) from host
 * Arrow buffers to host masks.  h_ref_* may be NULL when n_ref == 0.              */
int dyd_dedup_host(const int64_t* h_off, const uint8_t* h_bytes, const uint8_t* h_null, int64_t n,
                   int keep_mode, uint8_t* h_keep, int64_t* h_rep);
/* Anti-join (processor.py:194-199) from host Arrow buffers of both `source` columns to the host keep mask and the
 * first matching reference row (-1 when kept): H2D, hash of both columns, build + probe, D2H.           */
int dyd_antijoin_host(const int64_t* h_off, const uint8_t* h_bytes, const uint8_t* h_null, int64_t n,
                      const int64_t* h_ref_off, const uint8_t* h_ref_bytes, const uint8_t* h_ref_null, int64_t n_ref,
                      uint8_t* h_keep, int64_t* h_ref_row);

/* ------------------------------------------------ native ingest / egress (host) ---
 * Multi-threaded C++ replacement of the per-row json.loads / Python loops of processor.py:262-296
 * (step 4) and :341-366 (step 5).  `text`/`off` = the cells as one UTF-8 buffer with int64 offsets,
 * `is_text[r] == 0` marks non-string cells (NaN).  mode 0: polygons of every dict object (strict:
 * the text must be in json.dumps(ensure_ascii=False) form so that the output cell can be spliced);
 * mode 1: two-point boxes with the prefix-truncation rule.  Rows the parser is not certain about get
 * status 1 (SLOW) and contribute no objects: the caller handles them with CPython's json module.
 * The handle owns the parse results until dyd_ingest_free.                                        */
typedef struct dyd_ingest dyd_ingest;
int dyd_ingest_cells(const uint8_t* text, const int64_t* off, const uint8_t* is_text, int64_t n_rows,
                     int mode, int n_threads, dyd_ingest** out);
void dyd_ingest_free(dyd_ingest* h);
int dyd_ingest_sizes(const dyd_ingest* h, int64_t* n_obj, int64_t* n_vert, int64_t* n_slow);
/* Cells of modes 0 and 2 that are valid JSON in another style than json.dumps' are rewritten canonically
 * (= json.dumps(json.loads(text), ensure_ascii=False), same parsed value) and parsed from that form; every span
 * of such a row refers to the rewritten text.  n_canon = how many rows; new_off / new_text = the effective
 * texts (call with new_text == NULL for the offsets first) to pass to the export / egress calls instead of the
 * input.  dyd_json_canonical rewrites one document (test hook; -1: left to CPython, -2: buffer too small).   */
int dyd_ingest_effective_text(const dyd_ingest* h, const uint8_t* text, const int64_t* off, int64_t* n_canon,
                              int64_t* new_off, uint8_t* new_text, int n_threads);
int64_t dyd_json_canonical(const uint8_t* text, int64_t len, uint8_t* out, int64_t cap);
int dyd_ingest_export_polygons(const dyd_ingest* h, uint8_t* status, int64_t* img_off, int64_t* poly_off, double* xy,
                               int64_t* wh_off, int32_t* wh_len, uint8_t* wh_kind, int n_threads);
int dyd_ingest_export_boxes(const dyd_ingest* h, uint8_t* status, int64_t* img_off, double* pts, uint8_t* valid, int n_threads);
/* mode 2 (step 5.5, replace_labels_by_mapping processor.py:560-604): per dict object of "objects" the span
 * of its "name" string inside the cell's text (name_len -1: None / absent).  Only cells already in
 * json.dumps form are taken (status 0); status 4: "objects" absent or not a list (the reference leaves the
 * cell alone); status 1: the caller's CPython lane.  dyd_egress_names writes the new cell texts: the input
 * with the flagged objects' names replaced by JSON-escaped vocabulary entries.                        */
int dyd_ingest_export_names(const dyd_ingest* h, uint8_t* status, int64_t* cell_off, int64_t* name_off, int32_t* name_len,
                            int n_threads);
/* mode 2 for the split (step 6, processor.py:741-775): status 5 = "objects" present but not a list; list_len =
 * elements of the list (dicts or not); obj_off / obj_len = span of every dict object.  dyd_egress_split writes
 * the cells of the expanded rows: the document with "objects" replaced by one object renamed to one label.   */
int dyd_ingest_export_objects(const dyd_ingest* h, int32_t* list_len, int64_t* obj_off, int32_t* obj_len, int n_threads);
int dyd_egress_split(const dyd_ingest* h, const uint8_t* text, const int64_t* off, int64_t n_exp,
                     const int64_t* exp_cell, const int64_t* exp_obj, const int32_t* exp_tok,
                     const uint8_t* tok_bytes, const int64_t* tok_off, int64_t n_tok,
                     int64_t* out_off, uint8_t* out, int n_threads);
int dyd_egress_names(const dyd_ingest* h, const uint8_t* text, const int64_t* off, const uint8_t* obj_flag,
                     const int32_t* obj_new, const uint8_t* vocab_bytes, const int64_t* vocab_off, int64_t n_vocab,
                     int64_t* out_off, uint8_t* out, int n_threads);
/* Output cells of step 4: input text with every ptList value replaced by the two corner points built
 * from the original number literals selected by K1's arg indices.  Call with out == NULL to get the
 * offsets (out_off int64[n_rows+1]), then again with the buffer.                                  */
int dyd_egress_ptlist(const dyd_ingest* h, const uint8_t* text, const int64_t* off, const int32_t* arg,
                      const uint8_t* valid, int64_t* out_off, uint8_t* out, int n_threads);
/* Body rows of DataFrame.to_csv(index=False) for string / float64 / int64 / bool columns (kinds 0-3),
 * byte-identical to pandas + csv.QUOTE_MINIMAL.  out == NULL: fill row_off[n_rows+1]; else write. */
int dyd_csv_write(const int32_t* kinds, const int64_t* const* offs, const uint8_t* const* datas,
                  const uint8_t* const* valids, int32_t n_cols, int64_t n_rows, int64_t* row_off,
                  uint8_t* out, int n_threads);
/* DataFrame.to_csv(path, index=False) of the selected rows straight into the file (processor.py:158, 213, 310-313,
 * 402-407): `prefix` = BOM + header line (written first), rows = int64[n_sel] row numbers in output order (NULL: rows
 * 0..n_sel-1), columns as for dyd_csv_write.  Worker threads format blocks of rows while the calling thread writes the
 * finished blocks in order; neither the selected frame nor the whole body is materialised.  append != 0 opens with
 * O_APPEND instead of truncating.  Returns DYD_E_IO (errno set) when the file cannot be opened or written.        */
int dyd_csv_write_file(const char* path, int32_t append, const uint8_t* prefix, int64_t prefix_len,
                       const int32_t* kinds, const int64_t* const* offs, const uint8_t* const* datas,
                       const uint8_t* const* valids, int32_t n_cols, const int64_t* rows, int64_t n_sel,
                       int n_threads, int64_t* bytes_written);
int dyd_py_float_repr(double v, char* out40);     /* CPython repr(float); used by the canonical-form check */

/* YOLO label text of processor.py:1045-1052 from dyd_yolo_normalise's output: per image the lines
 * f"{cls} {cx:.6f} {cy:.6f} {bw:.6f} {bh:.6f}" of the boxes with ok != 0, joined by "\n" (host buffers).
 * out == NULL: fill out_off[n_img+1]; else write the text at those offsets.                        */
int dyd_yolo_format(const int64_t* img_off, const int32_t* class_id, const double* cxcywh, const uint8_t* ok,
                    int64_t n_img, int64_t* out_off, uint8_t* out, int n_threads);

/* ------------------------------------------------------- CSV ingest (§8f-2) ---
 * The tokenizer behind pd.read_csv(path, encoding="utf-8[-sig]") (processor.py:124-128, 181-182, 235,
 * 379, 424, 530, 678) for the text columns of the pipeline: pandas' C tokenizer state machine for the
 * default dialect, multi-threaded, producing Arrow large_string buffers.  Header names, dtype inference
 * of columns that are not certainly text, and every input outside the restated dialect (flags bit 0)
 * stay with pandas (deal_yolo_daya_b200/native.py:read_csv).
 * na_bytes/na_off: pandas' NA strings (pandas._libs.parsers.STR_NA_VALUES), packed.              */
int dyd_csv_open(const uint8_t* data, int64_t n, const uint8_t* na_bytes, const int64_t* na_off, int32_t n_na,
                 int32_t threads, void** handle);
int dyd_csv_info(void* handle, int64_t* n_rows, int32_t* n_cols, int64_t* header_begin, int64_t* header_end, int32_t* flags);
/* flags: bit 0 = outside the restated dialect (pandas reads the file), bit 1 = no record at all, bit 2 = the file was
 * tokenised by the AVX-512 scanner (statistic).                                                             */
/* window = rows per dtype-inference chunk of pandas' low-memory reader.  Per column: unescaped bytes,
 * missing cells, 1 iff every window holds a cell that is certainly text, 1 iff valid UTF-8.         */
int dyd_csv_measure(void* handle, int64_t window, int64_t* col_bytes, int64_t* col_nulls, uint8_t* col_text,
                    uint8_t* col_utf8, int32_t threads);
/* Arrow buffers of the selected columns: offsets int64[n_rows+1], data, validity bitmap (may be NULL). */
int dyd_csv_fill(void* handle, int32_t n_sel, const int32_t* cols, int64_t* const* off_out, uint8_t* const* data_out,
                 uint8_t* const* bitmap_out, int32_t threads);
void dyd_csv_close(void* handle);
/* Reads the first n bytes of the file into h_out with several threads (pread); DYD_E_IO + errno on failure.     */
int dyd_read_file(const char* path, uint8_t* h_out, int64_t n, int32_t threads);
/* 1 iff pd.read_csv would return this text column (Arrow large_string buffers; valid = one byte per row or NULL)
 * unchanged after DataFrame.to_csv wrote it: no valid cell empty / an NA string / holding NUL, and every window of
 * `window` rows holds a cell that is certainly text.  check_cells = 0: only the window rule (cells known clean).   */
int dyd_csv_roundtrip_check(const int64_t* off, const uint8_t* data, const uint8_t* valid, int64_t n_rows, int64_t window,
                            const uint8_t* na_bytes, const int64_t* na_off, int32_t n_na, int32_t check_cells, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* DYD_H_ */
