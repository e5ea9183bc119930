/* meshclust_b200 -- C-ABI of the B200-native MeShClust hot path.
 *
 * The reference (BioinformaticsToolsmith/MeShClust, C++11 + OpenMP) has no plugin / FFI interface:
 * its hot path is a set of C++ methods called from Runner / Trainer / ClusterFactory.  Each entry
 * point below replaces one of those call seams (cited as reference file:line); INTEGRATION.md
 * shows the binding a maintainer of the reference would add at each seam.
 *
 * Conventions
 *   - every call returns 0 on success or a negative MC_ERR_* code; mc_last_error() gives the text
 *     (thread-local).  No exception ever crosses this boundary.
 *   - all pointer arguments are HOST pointers unless the name ends in _dev.
 *   - a "row" is a point (one sequence) in the order the caller loaded it; the caller owns the
 *     mapping between rows and its own ids / headers.
 *   - histograms, sequences, alive flags, marks and the model stay resident in HBM between calls.
 *   - there is no CPU fallback: without a CUDA device every call except mc_last_error(),
 *     mc_version() and mc_host_segments() fails with MC_ERR_CUDA.
 */
#ifndef MESHCLUST_B200_H
#define MESHCLUST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_OK 0
#define MC_ERR_CUDA (-1)        /* CUDA runtime error / no device */
#define MC_ERR_ARG (-2)         /* bad argument */
#define MC_ERR_STATE (-3)       /* call order (e.g. scan before histograms are built) */
#define MC_ERR_INPUT (-4)       /* invalid nucleotide / sequence with no usable segment:
                                   the reference throws InvalidInputException (ChromosomeOneDigit.cpp:102-106)
                                   or std::out_of_range (Chromosome.cpp:193) */
#define MC_ERR_UNSUPPORTED (-5) /* k-mer count needs more than 16 bits, k too large, ... */

typedef struct mc_ctx mc_ctx;

const char *mc_version(void);
const char *mc_last_error(void);
int mc_device_count(void);

/* Replaces the OpenMP runtime set-up (Runner.cpp:201-214).  One context drives one GPU. */
int mc_ctx_create(mc_ctx **out, int device);
void mc_ctx_destroy(mc_ctx *ctx);
/* the cudaStream_t every kernel of this context is launched on (for external event timing) */
void *mc_stream(mc_ctx *ctx);
int mc_sync(mc_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py "gpu_launches") */
int64_t mc_launch_count(mc_ctx *ctx);

/* ---- stage 0: sequences ------------------------------------------------------------------ */

/* Host helper (pure C, no GPU): the non-N segment list of one raw sequence, exactly as
 * Chromosome::removeN / mergeSegments / makeSegmentList produce it (Chromosome.cpp:162-258).
 * Writes up to max_segs [start,end] pairs; returns the segment count, or -1 when the reference
 * would throw (no non-N run at all). */
int mc_host_segments(const uint8_t *letters, int64_t len, int32_t *segs, int max_segs);

/* Upload n raw sequences (letters as read from FASTA, any case, IUPAC allowed; concatenated,
 * offsets[n+1]) with their segment lists (seg_offsets[n+1] into segs[2*nseg]), and encode them
 * on the GPU into the reference's digit strings (ChromosomeOneDigit.cpp:95-144).
 * Replaces ChromosomeOneDigit::encodeNucleotides for every record of ChromListMaker's list. */
int mc_load_sequences(mc_ctx *ctx, const uint8_t *letters, const int64_t *offsets, int64_t n,
                      const int32_t *segs, const int64_t *seg_offsets);
/* FASTA ingest on the device.  Replaces, for files with LF line ends, the copy of every sequence
 * line into the record's string (ChromListMaker.cpp:92-120, Chromosome::appendToSequence, Chromosome.cpp:73-82)
 * and the scan for N of Chromosome::removeN (Chromosome.cpp:162-184).
 * raw[raw_bytes]: the files' bytes as read (host).  Record i (in the caller's ROW order) owns the bytes
 * [span_begin[i], span_end[i]) of raw -- its sequence lines, header line excluded -- which hold
 * offsets[i+1] - offsets[i] letters and otherwise only '\n'.  The letters land at offsets[i] of the device letter
 * buffer, exactly what mc_load_sequences would have uploaded.  rec_flags_out[i]: bit 0 = the record holds an N / n
 * (derive its segments with mc_host_segments), bit 1 = it holds a letter other than A C G T N in either case
 * (validate).  Must be followed by mc_load_segments.  MC_ERR_INPUT when spans and letter counts disagree. */
int mc_ingest_fasta(mc_ctx *ctx, const uint8_t *raw, int64_t raw_bytes, const int64_t *span_begin, const int64_t *span_end,
                    const int64_t *offsets, int64_t n, uint8_t *rec_flags_out);
/* Optional hint: make the context's scratch buffer at least `bytes` large now, so that later stages (ingest, distance
 * keys, alignments, mc_accumulate_run with its staging copies) do not have to replace it in the middle of a run. */
int mc_reserve_scratch(mc_ctx *ctx, int64_t bytes);
/* Optional: send the raw bytes ahead (blocking; meant for a helper thread) while the caller still derives the row
 * order and the spans.  An mc_ingest_fasta with the same raw / raw_bytes and at most n_records records that follows
 * without another call on ctx in between does not upload them again. */
int mc_stage_fasta_bytes(mc_ctx *ctx, const uint8_t *raw, int64_t raw_bytes, int64_t n_records);
/* The segment lists of the records ingested by mc_ingest_fasta (same layout as in mc_load_sequences).  validate != 0
 * checks every letter against the reference's code table (ChromosomeOneDigit.cpp:59-85; MC_ERR_INPUT = the
 * reference's InvalidInputException); callers pass 0 when no record has flag bit 1. */
int mc_load_segments(mc_ctx *ctx, const int32_t *segs, const int64_t *seg_offsets, int validate);
/* D2H copy of the letters as they stand on the device (before any alignment: the letters; afterwards the digit
 * strings) -- tests */
int mc_copy_letters(mc_ctx *ctx, uint8_t *out);
/* D2H copy of the encoded digit strings (same layout as the letters) -- tests / debugging */
int mc_copy_digits(mc_ctx *ctx, uint8_t *out);

/* ---- stage 1: k-mer histograms ----------------------------------------------------------- */

/* Build the dense 4^k count vector (pseudo-count 1) of every loaded sequence.
 * Replaces fill_table (ClusterFactory.h:40-55) -> KmerHashTable::wholesaleIncrement
 * (KmerHashTable.cpp:133-223) as called from Runner::run's pre-scan (Runner.cpp:57-67) and
 * ClusterFactory::get_divergence_point (ClusterFactory.cpp:989-1010).
 * tbytes = 1 or 2 selects uint8 / uint16 bins; 0 = choose like Runner.cpp:75-89 (smallest width
 * that holds the largest bin).  *tbytes_out / *max_count_out report what was used / found. */
int mc_build_histograms(mc_ctx *ctx, int k, int tbytes, int *tbytes_out, uint64_t *max_count_out);

/* Alternative to the two calls above: upload ready-made histograms (n x 4^k bins of tbytes) and
 * sequence lengths.  Point constants (mag, sum p^2) are computed on the GPU. */
int mc_load_histograms(mc_ctx *ctx, const void *hists, int tbytes, int k, const uint64_t *lens,
                       int64_t n);

int mc_copy_histograms(mc_ctx *ctx, void *out);
/* any of the three may be NULL */
int mc_copy_point_stats(mc_ctx *ctx, uint64_t *len, uint64_t *mag, uint64_t *sumsq);

/* ---- stage 2: pair features + GLM -------------------------------------------------------- */

/* Install the trained classifier: normalisation bounds in lookup order
 * [LD, INTERSECTION, MANHATTAN, PEARSON, KULCZYNSKI2] (Feature.cpp:15-28,87-114), GLM weights
 * w[0..nfeat] (Trainer.cpp:611-612), nfeat in {3,4} (Trainer.cpp:584-587,603-646).
 * Replaces the Feature<T> bounds + Trainer::weights that get_close/filter/merge read. */
int mc_set_model(mc_ctx *ctx, const double *mins, const double *maxs, const double *weights,
                 int nfeat);

/* DivergencePoint::distance (DivergencePoint.cpp:68-81) of every row against C center rows,
 * as used by the sort comparators of Trainer::split (Trainer.cpp:681-684,698-701).
 * keys_out[c*n + row], values in [0,10000]. */
int mc_distance_keys(mc_ctx *ctx, const int32_t *center_rows, int C, uint16_t *keys_out);

/* Raw features of m (a,b) row pairs: out5[m][5] = LD, INTERSECTION, MANHATTAN, PEARSON,
 * KULCZYNSKI2 (Feature.cpp:207-340), dist_out[m] = distance().  Either output may be NULL.
 * Replaces Feature::raw as driven by Feature::normalize (Feature.cpp:87-114). */
int mc_pair_features(mc_ctx *ctx, const int32_t *a, const int32_t *b, int64_t m, double *out5,
                     uint64_t *dist_out);

/* Classify m (a,b) row pairs with the installed model: GLM sum, f0 (first combo = "dist") and
 * flag = (round(1/(1+exp(-sum))) == 1).  Any output may be NULL.
 * Replaces Trainer::filter / Trainer::merge / generate_feat_mat loop bodies
 * (Trainer.cpp:334-349,129-157,367-414).  feats_out[m][4] (optional) gets the combo features. */
int mc_pair_classify(mc_ctx *ctx, const int32_t *a, const int32_t *b, int64_t m, double *sum_out,
                     double *f0_out, uint8_t *flag_out, double *feats_out);

/* Number of similar / not-similar decisions (mc_scan, mc_accumulate_step, mc_accumulate_run,
 * mc_update_centers, mc_pair_classify with flag_out) this context has taken on a GLM sum within
 * 1e-9 of the decision threshold (round(1/(1+exp(-sum))) == 1, Trainer.cpp:95,102) -- the only
 * pairs whose decision a differently rounded exp() could change.  reset != 0 zeroes the count. */
int mc_near_threshold_count(mc_ctx *ctx, int64_t *count_out, int reset);

typedef struct mc_scan_result {
	int64_t n_eval;   /* alive rows evaluated in [lo,hi] */
	int64_t n_pos;    /* rows classified similar (marked and removed from the alive set) */
	int64_t best_row; /* argmax of f0 over the evaluated rows, first maximum wins; -1 when no
	                     row has f0 > -1 (Trainer.cpp:42-48,99) */
	double best_f0;
} mc_scan_result;

/* Reset alive flags (all rows alive) and clear marks: a fresh bvec (Runner.cpp:342-350). */
int mc_alive_reset(mc_ctx *ctx);
/* Remove single rows from the alive set: bvec::pop / bvec::erase (bvec.cpp:27-38,281-285). */
int mc_alive_kill(mc_ctx *ctx, const int64_t *rows, int64_t m);

/* Trainer::get_close (Trainer.cpp:34-114) over the alive rows of the inclusive row range
 * [lo,hi], followed by bvec::remove_available (bvec.cpp:290-317): every row classified similar
 * to center_row is marked in marks_out[row-lo] (1 byte each, optional) and leaves the alive set. */
int mc_scan(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, mc_scan_result *res,
            uint8_t *marks_out);

/* Pipelined form of mc_scan for callers that keep several scans in flight (speculative seeds,
 * multi-center sweeps): mc_scan_enqueue launches the scan on the context's stream without waiting
 * and parks its summary in result slot `slot` (0 <= slot < MC_SCAN_SLOTS); remove_marked = 0
 * leaves the alive set untouched (get_close without the following remove_available).
 * mc_scan_collect waits for the stream and copies nslots summaries starting at slot0. */
#define MC_SCAN_SLOTS 1024
int mc_scan_enqueue(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked,
                    int slot);
int mc_scan_collect(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *res);
/* count scans enqueued back to back: scan i uses center_rows[i], [lo[i],hi[i]] and slot slot0+i.
 * remove_marked: MC_SCAN_KEEP (0) the scans are independent (several of them may share a launch);
 * MC_SCAN_REMOVE (1) every scan removes what it marks before the next one looks (accumulate()'s chain);
 * MC_SCAN_CHAIN (2) the scans run as such a dependent chain (one launch per scan, each behind the one
 * before it) but remove nothing -- the timing of MC_SCAN_REMOVE on an alive set that does not shrink. */
#define MC_SCAN_KEEP 0
#define MC_SCAN_REMOVE 1
#define MC_SCAN_CHAIN 2
int mc_scan_enqueue_many(mc_ctx *ctx, const int64_t *center_rows, const int64_t *lo,
                         const int64_t *hi, int count, int remove_marked, int slot0);

/* Device-side fold for sharded scans: reduces the partial records of nslots enqueued scans to one
 * mc_scan_result each, written to out_dev (DEVICE memory, nslots records) on the context's stream,
 * so a multi-GPU caller can hand them to a collective without a host round trip. */
int mc_scan_fold_dev(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *out_dev);
/* Adopt an externally owned cudaStream_t (e.g. the stream a communication library orders against)
 * for every later launch of this context; NULL restores the context's own stream. */
int mc_set_stream(mc_ctx *ctx, void *stream);

/* ---- multi-GPU: sharded scans ------------------------------------ */

/* `world` contexts, one per GPU (in one process or one process each), hold the same rows; the
 * scan work and the alive flags are sharded in blocks of consecutive rows (~256 KB of histograms:
 * 1024 rows at k = 4, 256 rows at k = 5): block b belongs to rank b mod world, so any length
 * window wider than a few blocks spreads over all GPUs and the owner of a row never changes.  What the reference reduces with OpenMP at the end of
 * Trainer::get_close (Trainer.cpp:38-48,81: arg-max, positives) crosses GPUs inside the scan
 * kernel: every CTA stores its partial into all ranks' inboxes over NVLink peer memory, and the
 * collect call folds world x SMs records on the device.
 *
 * mc_comm_init allocates this rank's inbox and writes its CUDA IPC handle to handle_out
 * (MC_COMM_HANDLE_BYTES, may be NULL for same-process use); mc_comm_connect takes the handles of
 * all ranks in rank order (exchanged by the caller, e.g. torch.distributed.all_gather_object);
 * mc_comm_connect_local wires contexts that live in one process.  At most 8 ranks. */
#define MC_COMM_HANDLE_BYTES 64
int mc_comm_init(mc_ctx *ctx, int rank, int world, uint8_t *handle_out);
int mc_comm_connect(mc_ctx *ctx, const uint8_t *handles);
int mc_comm_connect_local(mc_ctx *const *ctxs, int world);

/* get_close over this rank's tiles of [lo,hi], summary exchanged with all ranks.
 * Every rank must issue the same sequence of enqueue / collect calls (same center, range, slot);
 * a slot (0 <= slot < MC_XSLOTS) must be collected before it is enqueued again.  Collect returns
 * the same global summaries on every rank (rows are global row numbers); it fails with
 * MC_ERR_CUDA if a peer's records do not arrive within a few seconds. */
#define MC_XSLOTS 64
int mc_scan_sharded_enqueue(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi,
                            int remove_marked, int slot);
/* count scans back to back into slots slot0.. */
int mc_scan_sharded_enqueue_many(mc_ctx *ctx, const int64_t *center_rows, const int64_t *lo,
                                 const int64_t *hi, int count, int remove_marked, int slot0);
int mc_scan_sharded_collect(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *res);
/* collect = combine (device fold + copy to pinned host memory, asynchronous) + wait (host).  Callers
 * that keep the GPU busy enqueue the next scans between the two. */
int mc_scan_sharded_combine(mc_ctx *ctx, int slot0, int nslots);
int mc_scan_sharded_wait(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *res);
/* Streaming form for independent bursts of scans: enqueue this burst (scans on the context's
 * stream; fold + send + combine on a second, high-priority stream) and then wait for the summaries
 * of an earlier burst (slots prev_slot0.., e.g. the previous one); either count may be 0.  Bursts
 * should start at slots that are multiples of 16 (one completion event per bank). */
int mc_scan_sharded_burst(mc_ctx *ctx, const int64_t *center_rows, const int64_t *lo,
                          const int64_t *hi, int count, int remove_marked, int slot0, int prev_slot0,
                          int prev_count, mc_scan_result *prev_res);

/* ---- stage 3: mean-shift centers --------------------------------------------------------- */

/* get_mean (ClusterFactory.cpp:382-425): the member whose histogram is nearest
 * (DivergencePoint::distance_d, first minimum wins) to the mean of all members.
 * append != 0 extends the member list given by the previous call(s) instead of replacing it
 * (accumulate() grows `current`, ClusterFactory.cpp:689-692). */
int mc_mean_nearest(mc_ctx *ctx, const int64_t *rows, int64_t m, int append, int64_t *nearest_row,
                    double *nearest_dist);

typedef struct mc_step_result {
	mc_scan_result scan;  /* what mc_scan would have returned */
	int64_t nearest_row;  /* get_mean's pick after the marked rows joined `current`; -1 when scan.n_pos == 0 */
	int64_t n_members;    /* size of `current` after this step */
} mc_step_result;

/* One iteration of accumulate()'s inner loop (ClusterFactory.cpp:649-692) as a single submission
 * with a single host synchronisation: Trainer::get_close over the alive rows of [lo,hi]
 * (Trainer.cpp:34-114), bvec::remove_available (bvec.cpp:290-317) and -- when any row was marked --
 * get_mean (ClusterFactory.cpp:382-425) over `current` extended by the marked rows in iteration
 * order (ascending row).  restart != 0 begins a new cluster: `current` = {center_row}
 * (ClusterFactory.cpp:641).  The member list, its running bin sums and the marks never leave HBM;
 * marked_rows_out (optional, capacity cap >= number of marked rows, else MC_ERR_ARG) receives the
 * marked rows in ascending order so that the caller can drop them from its own bvec.
 * hi < lo is an empty range (no evaluation; nearest_row = -1). */
int mc_accumulate_step(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int restart,
                       mc_step_result *res, int64_t *marked_rows_out, int64_t cap);

typedef struct mc_run_stats {
	int64_t n_clusters;       /* centers pushed by accumulate() */
	int64_t n_scans;          /* get_close calls over a non-empty range */
	int64_t n_evals;          /* (center, point) pairs evaluated */
	int64_t n_near_threshold; /* of those, pairs whose GLM sum lies within 1e-9 of the decision threshold */
	int64_t n_steps;          /* iterations of accumulate()'s inner loop, empty ranges included */
	double device_seconds;    /* time the loop spent on the GPU */
	int64_t n_compactions;    /* times the rows still in the bvec were copied together (the scans stream those only) */
} mc_run_stats;

/* The whole of ClusterFactory::MS's first phase -- `while (last) accumulate(&last, points, ...)`
 * (ClusterFactory.cpp:722-729 with accumulate, :637-714) -- in one call and one kernel launch: the
 * bvec bookkeeping (bvec.cpp: pop :27-38, get_range :247-278 with index_of :123-149 and
 * inner_index_of :52-120, erase :281-285, remove_available :290-317), Trainer::get_close
 * (Trainer.cpp:34-114) and get_mean (ClusterFactory.cpp:382-425) run in a persistent device loop;
 * nothing returns to the host between two scans.
 * Rows must have been loaded in the bvec's iteration order: bin after bin, and inside a bin in the
 * order bvec::insert_finalize's sort left (non-decreasing length; checked, MC_ERR_ARG otherwise).
 * bin_bounds[nbins] are the bvec's bounds (bvec.cpp:10-24), bin_first_row[nbins + 1] the first row of
 * every bin (bin_first_row[nbins] = number of rows).  similarity is --id (the length window of a
 * center is [len * id, len / id], ClusterFactory.cpp:651-652).
 * Outputs (capacity: n rows each, cluster_offsets_out n + 1): cluster c has center row
 * center_rows_out[c] and the members member_rows_out[cluster_offsets_out[c] ..
 * cluster_offsets_out[c + 1]) in the order accumulate() collected them (seed first).
 * The run starts from a fresh bvec (all rows alive) and keeps its own alive set: the alive flags
 * mc_scan / mc_accumulate_step use are neither read nor written.
 * MC_ERR_UNSUPPORTED for histogram rows the staged scan kernel does not handle (k = 1, k >= 7) or a
 * bvec with more bins than one SM's shared memory holds (~5000): use mc_accumulate_step then. */
int mc_accumulate_run(mc_ctx *ctx, double similarity, const uint64_t *bin_bounds,
                      const int64_t *bin_first_row, int64_t nbins, int64_t *center_rows_out,
                      int64_t *cluster_offsets_out, int64_t *member_rows_out, mc_run_stats *stats);

/* Copy histograms, point constants, alive flags and the model of `src` into `dst`, a context on
 * another GPU of the same process (device-to-device over NVLink): the one-time replication that
 * stands in for the per-scan center broadcast of SURVEY 8(e). */
int mc_clone_points(mc_ctx *dst, mc_ctx *src);
/* The same for the sequences (letters / digit strings, offsets, segments): K4 is split by pairs over the GPUs
 * (Trainer.cpp:253-333, :703-721), and each of them needs both strings of its pairs. */
int mc_clone_sequences(mc_ctx *dst, mc_ctx *src);

/* mc_accumulate_step across `world` contexts of one process wired with mc_comm_connect_local:
 * rank r evaluates its tiles of the range and owns their alive flags, the marks of
 * all ranks land in rank 0's array over peer memory, rank 0 waits for every rank's summaries on the
 * device and runs the tail (compaction, running sums, get_mean).  Same results, same arguments. */
int mc_accumulate_step_sharded(mc_ctx *const *ctxs, int world, int64_t center_row, int64_t lo,
                               int64_t hi, int restart, mc_step_result *res,
                               int64_t *marked_rows_out, int64_t cap);

/* Re-number the first `count` rows: new row i (i < count) takes the histogram and constants of old
 * row old_of_new[i] (a permutation of 0..count-1); rows below n_alive are flagged alive, rows in
 * [n_alive, count) dead; rows >= count are untouched.  The host uses it while accumulate()
 * (ClusterFactory.cpp:637-714) consumes the points: rows that already belong to a cluster are moved
 * behind the alive ones (keeping the iteration order of both groups), so the scans that follow
 * stream alive rows only.  Nothing the reference computes depends on the numbering.  Sequences keep
 * their original rows: mc_align_pairs returns MC_ERR_STATE after a permutation.  The member list of
 * mc_accumulate_step / mc_mean_nearest must be empty or about to be restarted. */
int mc_permute_rows(mc_ctx *ctx, const int64_t *old_of_new, int64_t count, int64_t n_alive);
/* optional: allocate mc_permute_rows' staging buffers for all rows ahead of time */
int mc_reserve_permute(mc_ctx *ctx);

/* One Jacobi sweep of mean_shift_update (ClusterFactory.cpp:289-380) for ncenters centers:
 * center c sees the candidate rows cand_rows[cand_begin[c] .. cand_end[c]) (members of clusters
 * c-delta..c+delta in order), keeps those Trainer::filter (Trainer.cpp:334-349) classifies
 * similar to center_rows[c], averages them and returns in next_rows[c] the survivor nearest to
 * that mean (Trainer::closest, Trainer.cpp:351-365), or -1 when none survives. */
int mc_update_centers(mc_ctx *ctx, const int64_t *center_rows, int64_t ncenters,
                      const int64_t *cand_rows, int64_t ncand, const int64_t *cand_begin,
                      const int64_t *cand_end, int64_t *next_rows);

/* ---- stage 4: global alignment identity -------------------------------------------------- */

/* GlobAlignE(seq_a, seq_b, match 1, mismatch -1, open 2, continue 1) for m row pairs
 * (GlobAlignE.cpp:123-305 via Trainer::align, Trainer.cpp:15-31, and Feature::align,
 * Feature.cpp:222-243).  a is the reference's seq1, b its seq2 (results depend on the order in
 * corner cases).  identity = matches / alen is left to the caller (0/0 = NaN as in the reference). */
int mc_align_pairs(mc_ctx *ctx, const int32_t *a, const int32_t *b, int64_t m, int32_t *score,
                   int32_t *alen, int32_t *matches);

/* ---- host-buffer one-shot entry points (end-to-end measurement, simple embedding) -------- */

/* letters -> histograms in one call: upload, encode, count, download.  hists_out: n x 4^k bins. */
int mc_kmer_histograms_host(mc_ctx *ctx, const uint8_t *letters, const int64_t *offsets, int64_t n,
                            int k, int tbytes, void *hists_out, uint64_t *max_count_out);

/* get_close for ncenters centers over host histograms: uploads the n x 4^k histograms + lengths,
 * evaluates every (center, row) pair with the installed model and returns, per center, the mark
 * bytes (marks_out[c*n + row], optional) and the scan summary.  Rows are not removed between
 * centers (each center sees all n rows). */
int mc_scan_host(mc_ctx *ctx, const void *hists, int tbytes, int k, const uint64_t *lens, int64_t n,
                 const int64_t *center_rows, int ncenters, mc_scan_result *res, uint8_t *marks_out);

#ifdef __cplusplus
}
#endif
#endif /* MESHCLUST_B200_H */
