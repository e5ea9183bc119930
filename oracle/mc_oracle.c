/* TEST INFRASTRUCTURE ONLY -- see mc_oracle.h.  Plain-C restatement of the reference hot path.
 * Built by oracle/Makefile with -ffp-contract=off: the two places where the compiled reference
 * (-O3 -march=native, GCC default -ffp-contract=fast) fuses a multiply-add are written as explicit
 * fma() calls below, so this file gives the same bits on any host.
 */
#define _GNU_SOURCE
#include "mc_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * FASTA record body -> digit string + segments
 * ------------------------------------------------------------------------------------------- */

/* nonltr/ChromosomeOneDigit.cpp:59-85: A0 C1 G2 T3, IUPAC collapsed, N->C, X->G; others invalid */
static int code_of(int c) {
	switch (c) {
	case 'A': case 'M': case 'V': return 0;
	case 'C': case 'Y': case 'H': case 'N': return 1;
	case 'G': case 'R': case 'S': case 'X': return 2;
	case 'T': case 'K': case 'W': case 'B': case 'D': return 3;
	default: return -1;
	}
}

#define SEG_SPLIT 1000000   /* Chromosome.cpp:104 help(1000000, true) */

int mco_encode(const char *seq, long len, char *digits, int *segs, int max_segs) {
	long i;
	/* Chromosome.cpp:153-157 toUpperCase */
	for (i = 0; i < len; i++) digits[i] = (char)toupper((unsigned char)seq[i]);

	/* Chromosome.cpp:162-184 removeN: maximal non-N runs.  Quirk kept: a run that STARTS on the
	 * very last character is never closed (the "last index" branch is an else-if of the
	 * "run starts" branch), so it yields no segment. */
	long cap = 16, nraw = 0;
	long *raw = (long *)malloc(sizeof(long) * 2 * cap);
	long start = -1;
	for (i = 0; i < len; i++) {
		int isn = digits[i] == 'N';
		if (!isn && start == -1) {
			start = i;
		} else if (isn && start != -1) {
			if (nraw == cap) { cap *= 2; raw = (long *)realloc(raw, sizeof(long) * 2 * cap); }
			raw[2 * nraw] = start; raw[2 * nraw + 1] = i - 1; nraw++;
			start = -1;
		} else if (i == len - 1 && !isn && start != -1) {
			if (nraw == cap) { cap *= 2; raw = (long *)realloc(raw, sizeof(long) * 2 * cap); }
			raw[2 * nraw] = start; raw[2 * nraw + 1] = i; nraw++;
			start = -1;
		}
	}
	if (nraw == 0) { free(raw); return -1; }   /* segment->at(0) throws, Chromosome.cpp:193 */

	/* Chromosome.cpp:190-226 mergeSegments: join when next.start - cur.end < 10, keep >= 20 bp */
	long nm = 0;
	long *mer = (long *)malloc(sizeof(long) * 2 * nraw);
	long s = raw[0], e = raw[1];
	for (i = 1; i < nraw; i++) {
		long s1 = raw[2 * i], e1 = raw[2 * i + 1];
		if (s1 - e < 10) {
			e = e1;
		} else {
			if (e - s + 1 >= 20) { mer[2 * nm] = s; mer[2 * nm + 1] = e; nm++; }
			s = s1; e = e1;
		}
	}
	if (e - s + 1 >= 20) { mer[2 * nm] = s; mer[2 * nm + 1] = e; nm++; }
	free(raw);

	/* Chromosome.cpp:228-258 makeSegmentList: cut segments longer than 1 Mbp into floor(len/1M)
	 * pieces, the last one taking the remainder */
	int nseg = 0;
	for (i = 0; i < nm; i++) {
		long ss = mer[2 * i], ee = mer[2 * i + 1];
		long l = ee - ss + 1;
		if (l > SEG_SPLIT) {
			long frag = l / SEG_SPLIT, h;
			for (h = 0; h < frag; h++) {
				long fs = ss + h * SEG_SPLIT;
				long fe = (h == frag - 1) ? ee : fs + SEG_SPLIT - 1;
				if (nseg < max_segs) { segs[2 * nseg] = (int)fs; segs[2 * nseg + 1] = (int)fe; }
				nseg++;
			}
		} else {
			if (nseg < max_segs) { segs[2 * nseg] = (int)ss; segs[2 * nseg + 1] = (int)ee; }
			nseg++;
		}
	}

	/* ChromosomeOneDigit.cpp:95-144 encodeNucleotides: inside segments every letter (N included,
	 * as C) becomes a digit; outside, N stays the byte 'N' and the rest become digits.  With no
	 * segment at all nothing is touched (the "skipped segments" pass is guarded by segNum > 0). */
	if (nseg > 0) {
		long si = 0;
		for (i = 0; i < len; i++) {
			int inside = 0;
			while (si < nm && mer[2 * si + 1] < i) si++;
			if (si < nm && mer[2 * si] <= i) inside = 1;
			if (!inside && digits[i] == 'N') continue;
			int c = code_of((unsigned char)digits[i]);
			if (c < 0) { free(mer); return -1; }   /* InvalidInputException */
			digits[i] = (char)c;
		}
	}
	free(mer);
	return nseg;
}

/* ClusterFactory.h:40-55 fill_table + KmerHashTable.cpp:107-159,194-207 with init value 1
 * (ClusterFactory.cpp:995): for each segment [s,e] every start in [s, e-k+1] adds one to the bin
 * whose index is the base-4 big-endian value of the k digits. */
void mco_hist_digits(const char *digits, const int *segs, int nseg, int k, uint64_t *out) {
	const uint64_t nb = (uint64_t)1 << (2 * k);
	uint64_t b;
	int g;
	for (b = 0; b < nb; b++) out[b] = 1;
	for (g = 0; g < nseg; g++) {
		long s = segs[2 * g], last = (long)segs[2 * g + 1] - k + 1, i;
		/* (segments are >= 20 bp, so last >= s for every k whose 4^k table is buildable) */
		for (i = s; i <= last; i++) {
			uint64_t idx = 0;
			int j;
			for (j = 0; j < k; j++) idx = idx * 4 + (uint64_t)(unsigned char)digits[i + j];
			out[idx & (nb - 1)]++;
		}
	}
}

long mco_hist_batch(const char *seqs, const int64_t *offs, int n, int k, int tbytes, void *out,
                    uint64_t *max_count) {
	const uint64_t nb = (uint64_t)1 << (2 * k);
	long bad = 0;
	uint64_t gmax = 0;
	int i;
#pragma omp parallel for schedule(dynamic, 16) reduction(max : gmax)
	for (i = 0; i < n; i++) {
		long len = (long)(offs[i + 1] - offs[i]);
		char *dig = (char *)malloc((size_t)len + 1);
		int cap = 64;
		int *segs = (int *)malloc(sizeof(int) * 2 * cap);
		int ns = mco_encode(seqs + offs[i], len, dig, segs, cap);
		if (ns > cap) {
			cap = ns;
			segs = (int *)realloc(segs, sizeof(int) * 2 * cap);
			ns = mco_encode(seqs + offs[i], len, dig, segs, cap);
		}
		if (ns < 0) {
#pragma omp critical
			if (bad == 0 || -(long)(i + 1) > bad) bad = -(long)(i + 1);
		} else {
			uint64_t *h = (uint64_t *)malloc(sizeof(uint64_t) * nb);
			uint64_t b;
			mco_hist_digits(dig, segs, ns, k, h);
			for (b = 0; b < nb; b++) {
				if (h[b] > gmax) gmax = h[b];
				switch (tbytes) {
				case 1: ((uint8_t *)out)[(uint64_t)i * nb + b] = (uint8_t)h[b]; break;
				case 2: ((uint16_t *)out)[(uint64_t)i * nb + b] = (uint16_t)h[b]; break;
				case 4: ((uint32_t *)out)[(uint64_t)i * nb + b] = (uint32_t)h[b]; break;
				default: ((uint64_t *)out)[(uint64_t)i * nb + b] = h[b]; break;
				}
			}
			free(h);
		}
		free(dig);
		free(segs);
	}
	if (max_count) *max_count = gmax;
	return bad;
}

/* ---------------------------------------------------------------------------------------------
 * pair arithmetic
 * ------------------------------------------------------------------------------------------- */

static inline uint64_t bin_at(const void *h, int tbytes, int i) {
	switch (tbytes) {
	case 1: return ((const uint8_t *)h)[i];
	case 2: return ((const uint16_t *)h)[i];
	case 4: return ((const uint32_t *)h)[i];
	default: return ((const uint64_t *)h)[i];
	}
}

void mco_pair_stats(const void *p, const void *q, int nbins, int tbytes, uint64_t *summin,
                    uint64_t *dot) {
	uint64_t s = 0, d = 0;
	int i;
	for (i = 0; i < nbins; i++) {
		uint64_t a = bin_at(p, tbytes, i), b = bin_at(q, tbytes, i);
		s += a < b ? a : b;
		d += a * b;
	}
	*summin = s;
	*dot = d;
}

void mco_point_stats(const void *p, int nbins, int tbytes, uint64_t *mag, uint64_t *sq) {
	uint64_t m = 0, s = 0;
	int i;
	for (i = 0; i < nbins; i++) {
		uint64_t a = bin_at(p, tbytes, i);
		m += a;
		s += a * a;
	}
	*mag = m;
	*sq = s;
}

/* DivergencePoint.cpp:68-81: (uint64)(10000 * (1 - f*f)), f = 2*summin / (mag_p + mag_q).
 * The compiled reference evaluates 1 - f*f as one fused negate-multiply-add. */
static uint64_t distance_from(uint64_t summin, uint64_t magsum) {
	double frac = (double)(2 * summin) / (double)magsum;
	return (uint64_t)(10000.0 * fma(-frac, frac, 1.0));
}

void mco_features(const void *p, const void *q, int nbins, int tbytes, uint64_t lp, uint64_t lq,
                  double *out5, uint64_t *dist) {
	uint64_t S, D, mp, mq, sp, sq;
	const int N = nbins;
	mco_pair_stats(p, q, nbins, tbytes, &S, &D);
	mco_point_stats(p, nbins, tbytes, &mp, &sp);
	mco_point_stats(q, nbins, tbytes, &mq, &sq);

	/* Feature.cpp:326-339 length_difference */
	out5[0] = (double)(lp > lq ? lp - lq : lq - lp);
	/* Feature.cpp:259-271 intersection */
	out5[1] = (double)(2 * S) / (double)(mp + mq);
	/* Feature.cpp:311-323 manhattan: sum |p-q| accumulated in an int = mag_p + mag_q - 2 summin */
	out5[2] = (double)(int)(mp + mq - 2 * S);
	/* Feature.cpp:274-294 pearson around the ROUNDED integer means */
	{
		double dap = (double)mp / N, daq = (double)mq / N;
		int64_t ap = (int)round(dap), aq = (int)round(daq);
		int64_t np = (int64_t)sp - 2 * ap * (int64_t)mp + (int64_t)N * ap * ap;
		int64_t nq = (int64_t)sq - 2 * aq * (int64_t)mq + (int64_t)N * aq * aq;
		int64_t dot = (int64_t)D - aq * (int64_t)mp - ap * (int64_t)mq + (int64_t)N * ap * aq;
		double prod = (double)(np * nq);
		out5[3] = (double)dot / sqrt(prod > 0.5 ? prod : 0.5);
	}
	/* Feature.cpp:207-219 kulczynski2 with the UNROUNDED means */
	{
		double ap = (double)mp / N, aq = (double)mq / N;
		double coeff = N * (ap + aq) / (2 * ap * aq);
		out5[4] = coeff * (double)S;
	}
	if (dist) *dist = distance_from(S, mp + mq);
}

/* DivergencePoint.cpp:53-65: the mean is truncated to T inside the min; mag is a uint64 that is
 * bumped by a double each step (so it advances by floor(p_i + c_i)). */
double mco_distance_d(const void *p, int nbins, int tbytes, const double *mean) {
	uint64_t dist = 0, mag = 0;
	int i;
	for (i = 0; i < nbins; i++) {
		uint64_t a = bin_at(p, tbytes, i), c;
		switch (tbytes) {
		case 1: c = (uint8_t)mean[i]; break;
		case 2: c = (uint16_t)mean[i]; break;
		case 4: c = (uint32_t)mean[i]; break;
		default: c = (uint64_t)mean[i]; break;
		}
		dist += 2 * (a < c ? a : c);
		mag = (uint64_t)((double)mag + ((double)a + mean[i]));
	}
	{
		double frac = (double)dist / (double)mag;
		return 10000.0 * fma(-frac, frac, 1.0);
	}
}

/* ClusterFactory.cpp:316-335,395-402 + DivergencePoint.cpp:155-173: sum as doubles (exact for
 * integers), then one divide per bin */
void mco_mean(const void *hists, int nbins, int tbytes, int m, double *mean) {
	int i, j;
	for (i = 0; i < nbins; i++) mean[i] = 0;
	for (j = 0; j < m; j++) {
		const char *h = (const char *)hists + (size_t)j * nbins * tbytes;
		for (i = 0; i < nbins; i++) mean[i] += (double)bin_at(h, tbytes, i);
	}
	for (i = 0; i < nbins; i++) mean[i] /= (double)m;
}

/* Trainer.cpp:81-106 (get_close body) on top of Feature.h:64-88 / Feature.cpp:42-51.
 * lookup order [LD, INTERSECTION, MANHATTAN, PEARSON, KULCZYNSKI2]; LD, MANHATTAN, PEARSON are
 * distances (1 - v'), the others similarities; combos (Trainer.cpp:584-587):
 * f0 = LD'*INT', f1 = LD'^2*MAN'^2, f2 = PEARSON', f3 = LD'^2*KUL'^2. */
void mco_scan(const void *hists, const uint64_t *lens, int n, int nbins, int tbytes,
              const void *center, uint64_t center_len, const double *mins, const double *maxs,
              const double *weights, int nfeat, double *sum_out, double *f0_out,
              uint8_t *flag_out) {
	static const int is_sim[5] = {0, 1, 0, 0, 1};
	const int nlookup = nfeat >= 4 ? 5 : 4;
	int i;
#pragma omp parallel for schedule(static)
	for (i = 0; i < n; i++) {
		const char *h = (const char *)hists + (size_t)i * nbins * tbytes;
		double c[5], f[4], sum;
		int j;
		mco_features(h, center, nbins, tbytes, lens[i], center_len, c, NULL);
		for (j = 0; j < nlookup; j++) {
			double v = (c[j] - mins[j]) / (maxs[j] - mins[j]);
			c[j] = is_sim[j] ? v : 1 - v;
		}
		f[0] = (1.0 * c[0]) * c[1];
		f[1] = (1.0 * (c[0] * c[0])) * (c[2] * c[2]);
		f[2] = 1.0 * c[3];
		f[3] = nfeat >= 4 ? (1.0 * (c[0] * c[0])) * (c[4] * c[4]) : 0.0;
		sum = weights[0];
		for (j = 0; j < nfeat; j++) sum = fma(weights[j + 1], f[j], sum);
		sum_out[i] = sum;
		f0_out[i] = f[0];
		flag_out[i] = (round(1.0 / (1 + exp(-sum))) == 1.0);
	}
}

/* ---------------------------------------------------------------------------------------------
 * GlobAlignE (utility/GlobAlignE.cpp:123-292): Gotoh global alignment that carries the length of
 * the alignment and the number of matching columns along the winning path.  Restated with
 * explicit previous/current rows of {score,len,id} triples instead of the reference's nine
 * rolling arrays + lag scalars.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int sc, len, id; } cell_t;

void mco_globalign(const char *s1, int la, const char *s2, int lb, int match, int mismatch,
                   int gopen, int gcont, int *score, int *alen, int *matches) {
	const int shorter = la < lb ? la : lb;
	const int diff = la > lb ? la - lb : lb - la;
	/* :125-135 finite "minus infinity" that takes part in the arithmetic */
	const int ninf = (diff >= 1 ? -gopen - diff * gcont : 0) + mismatch * shorter - 1;
	cell_t *M = (cell_t *)malloc(sizeof(cell_t) * (size_t)(la + 1) * 2);
	cell_t *U = (cell_t *)malloc(sizeof(cell_t) * (size_t)(la + 1) * 2);
	cell_t *L = (cell_t *)malloc(sizeof(cell_t) * (size_t)(la + 1) * 2);
	cell_t *Mp = M, *Mc = M + la + 1, *Up = U, *Uc = U + la + 1, *Lp = L, *Lc = L + la + 1, *t;
	int i, j;

	/* row 0 (:137-160) */
	for (i = 0; i <= la; i++) {
		Mp[i].sc = i == 0 ? 0 : ninf; Mp[i].len = i; Mp[i].id = 0;
		Up[i].sc = ninf;              Up[i].len = i; Up[i].id = 0;
		Lp[i].sc = i == 0 ? ninf : -gopen - i * gcont; Lp[i].len = i; Lp[i].id = 0;
	}
	for (j = 1; j <= lb; j++) {
		/* diagonal predecessors of column 1 live in column 0 of the previous row; the U one
		 * is synthesised (:164-170) */
		cell_t dM = Mp[0], dL = Lp[0], dU;
		dU.sc = -gopen - (j - 1) * gcont; dU.len = j - 1; dU.id = 0;
		for (i = 1; i <= la; i++) {
			/* vertical gap (:178-193): open from M wins ties */
			int ob = Mp[i].sc - (gopen + gcont), oc = Up[i].sc - gcont;
			int s, m, x, y, best;
			cell_t src;
			if (ob >= oc) { Uc[i].sc = ob; Uc[i].len = Mp[i].len + 1; Uc[i].id = Mp[i].id; }
			else          { Uc[i].sc = oc; Uc[i].len = Up[i].len + 1; Uc[i].id = Up[i].id; }
			/* diagonal (:201-241): tie order M, L, U */
			s = (s1[i - 1] == s2[j - 1]) ? match : mismatch;
			m = dM.sc + s; x = dL.sc + s; y = dU.sc + s;
			best = m > x ? m : x; if (y > best) best = y;
			src = (best == m) ? dM : (best == x) ? dL : dU;
			Mc[i].sc = best; Mc[i].len = src.len + 1; Mc[i].id = src.id + (s == match ? 1 : 0);
			dM = Mp[i]; dL = Lp[i]; dU = Up[i];
		}
		/* column 0 of this row (:250-256); U[0] is never rewritten */
		Mc[0].sc = ninf; Mc[0].len = j; Mc[0].id = 0;
		Lc[0].sc = ninf; Lc[0].len = j; Lc[0].id = 0;
		Uc[0] = Up[0];
		/* horizontal gap on the CURRENT row (:258-273): open from M wins ties */
		for (i = 1; i <= la; i++) {
			int ob = Mc[i - 1].sc - (gopen + gcont), oc = Lc[i - 1].sc - gcont;
			if (ob >= oc) { Lc[i].sc = ob; Lc[i].len = Mc[i - 1].len + 1; Lc[i].id = Mc[i - 1].id; }
			else          { Lc[i].sc = oc; Lc[i].len = Lc[i - 1].len + 1; Lc[i].id = Lc[i - 1].id; }
		}
		t = Mp; Mp = Mc; Mc = t;
		t = Up; Up = Uc; Uc = t;
		t = Lp; Lp = Lc; Lc = t;
	}
	/* :278-291 tie order M, L, U */
	{
		int best = Mp[la].sc > Lp[la].sc ? Mp[la].sc : Lp[la].sc;
		cell_t w;
		if (Up[la].sc > best) best = Up[la].sc;
		w = (best == Mp[la].sc) ? Mp[la] : (best == Lp[la].sc) ? Lp[la] : Up[la];
		*score = best; *alen = w.len; *matches = w.id;
	}
	free(M); free(U); free(L);
}

void mco_globalign_batch(const char *seqs, const int64_t *offs, const int32_t *pa,
                         const int32_t *pb, int npairs, int *score, int *alen, int *matches) {
	int i;
#pragma omp parallel for schedule(dynamic)
	for (i = 0; i < npairs; i++) {
		int a = pa[i], b = pb[i];
		/* Trainer.cpp:25-27 / Feature.cpp:233-235: always (1, -1, 2, 1) */
		mco_globalign(seqs + offs[a], (int)(offs[a + 1] - offs[a]), seqs + offs[b],
		              (int)(offs[b + 1] - offs[b]), 1, -1, 2, 1, score + i, alen + i, matches + i);
	}
}
