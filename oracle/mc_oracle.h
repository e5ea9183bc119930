/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for the MeShClust hot path.
 *
 * Plain-C restatement of the reference algorithm (each function cites the reference file:line it
 * follows).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product (meshclust_b200/csrc, the
 * C-ABI in include/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function below against
 * the compiled, unmodified reference (oracle/_ref/libmcref.so, built by oracle/Makefile from
 * /root/reference) on seeded inputs, and against the known-answer vectors of SURVEY.md App. B
 * (tests/golden/), which were produced by the reference code itself.
 */
#ifndef MC_ORACLE_H
#define MC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* letters -> digit string + segment list.  returns #segments, -1 = reference would throw. */
int mco_encode(const char *seq, long len, char *digits, int *segs, int max_segs);

/* dense 4^k histogram with pseudo-count 1 over the segments of one encoded sequence. */
void mco_hist_digits(const char *digits, const int *segs, int nseg, int k, uint64_t *out);

/* batch: raw letters -> n x 4^k bins of tbytes (1,2,4,8) width; *max_count = largest bin.
 * returns 0, or -(i+1) when sequence i is invalid. */
long mco_hist_batch(const char *seqs, const int64_t *offs, int n, int k, int tbytes, void *out,
                    uint64_t *max_count);

/* the two pairwise reductions everything else derives from */
void mco_pair_stats(const void *p, const void *q, int nbins, int tbytes, uint64_t *summin,
                    uint64_t *dot);

/* per-point constants: mag = sum p, sq = sum p^2 */
void mco_point_stats(const void *p, int nbins, int tbytes, uint64_t *mag, uint64_t *sq);

/* raw features [LD, INTERSECTION, MANHATTAN, PEARSON, KULCZYNSKI2] and DivergencePoint::distance */
void mco_features(const void *p, const void *q, int nbins, int tbytes, uint64_t lp, uint64_t lq,
                  double *out5, uint64_t *dist);

/* DivergencePoint::distance_d against a double mean */
double mco_distance_d(const void *p, int nbins, int tbytes, const double *mean);

/* mean of m histograms (exact integer sum, one divide per bin) */
void mco_mean(const void *hists, int nbins, int tbytes, int m, double *mean);

/* get_close-style evaluation of n points against one center with a trained model */
void mco_scan(const void *hists, const uint64_t *lens, int n, int nbins, int tbytes,
              const void *center, uint64_t center_len, const double *mins, const double *maxs,
              const double *weights, int nfeat, double *sum_out, double *f0_out,
              uint8_t *flag_out);

/* GlobAlignE: score, alignment length, matches on the chosen path */
void mco_globalign(const char *s1, int la, const char *s2, int lb, int match, int mismatch,
                   int gopen, int gcont, int *score, int *alen, int *matches);

void mco_globalign_batch(const char *seqs, const int64_t *offs, const int32_t *pa,
                         const int32_t *pb, int npairs, int *score, int *alen, int *matches);

#ifdef __cplusplus
}
#endif
#endif
