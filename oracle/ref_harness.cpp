// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Unit-level oracle harness: thin extern "C" shims (this file is ours) around
// the UNMODIFIED reference sources, which oracle/Makefile compiles from where
// they lie under /root/reference/src into oracle/_ref/libmcref.so.  Used by
// tests/ (through ctypes) to pin oracle/mc_oracle.c and the CUDA path to the
// reference's own arithmetic, and by bench.py's reference arm.
//
// Each shim names the reference entry point it drives.
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <stdexcept>
// every std header the reference pulls in must be seen BEFORE the access hack below
#include <algorithm>
#include <cmath>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <random>
#include <set>
#include <sstream>
#include <tuple>
#include <cfenv>
#include <cstdlib>
#include <cstdio>
#include <ctime>
#include <iomanip>
#include <utility>
#include <queue>
#include <list>
#include <cassert>
#include <dirent.h>
#include <sys/stat.h>

#define private public
#define protected public
#include "cluster/src/DivergencePoint.h"
#include "cluster/src/Feature.h"
#include "cluster/src/ClusterFactory.h"
#include "utility/GlobAlignE.h"
#include "nonltr/ChromosomeOneDigit.h"
#include "nonltr/KmerHashTable.h"
#undef private
#undef protected

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace nonltr;

namespace {

template <class T>
DivergencePoint<T> *mk_point(const void *hist, int nbins, uint64_t len) {
	const T *h = static_cast<const T *>(hist);
	std::vector<T> v(h, h + nbins);
	auto *p = new DivergencePoint<T>(v, len);   // DivergencePoint.cpp:97-109
	p->set_length(len);
	return p;
}

template <class T>
void features_t(const void *ph, const void *qh, int nbins, uint64_t lp, uint64_t lq,
		double *out5, uint64_t *dist) {
	DivergencePoint<T> *p = mk_point<T>(ph, nbins, lp);
	DivergencePoint<T> *q = mk_point<T>(qh, nbins, lq);
	out5[0] = Feature<T>::length_difference(*p, *q);   // Feature.cpp:326-339
	out5[1] = Feature<T>::intersection(*p, *q);        // Feature.cpp:259-271
	out5[2] = Feature<T>::manhattan(*p, *q);           // Feature.cpp:311-323
	out5[3] = Feature<T>::pearson(*p, *q);             // Feature.cpp:274-294
	out5[4] = Feature<T>::kulczynski2(*p, *q);         // Feature.cpp:207-219
	*dist = p->distance(*q);                           // DivergencePoint.cpp:68-81
	delete p;
	delete q;
}

template <class T>
double distance_d_t(const void *ph, int nbins, const double *mean) {
	DivergencePoint<T> *p = mk_point<T>(ph, nbins, 1);
	std::vector<double> m(mean, mean + nbins);
	DivergencePoint<double> c(m, 1);
	double d = p->distance_d(c);                       // DivergencePoint.cpp:53-65
	delete p;
	return d;
}

// mean exactly as get_mean / mean_shift_update build it
// (ClusterFactory.cpp:382-403 / :316-335): set_arg_to_this_d + operator+= + operator/=
template <class T>
void mean_t(const void *hists, int nbins, int m, double *mean_out) {
	const T *h = static_cast<const T *>(hists);
	DivergencePoint<T> *first = mk_point<T>(h, nbins, 1);
	Point<double> *top = first->create_double();
	top->zero();
	Point<double> *temp = top->clone();
	for (int i = 0; i < m; i++) {
		DivergencePoint<T> *p = mk_point<T>(h + (size_t)i * nbins, nbins, 1);
		p->set_arg_to_this_d(*temp);
		*top += *temp;
		delete p;
	}
	*top /= (double)m;
	auto &v = dynamic_cast<DivergencePoint<double> *>(top)->points;
	for (int i = 0; i < nbins; i++) mean_out[i] = v[i];
	delete top;
	delete temp;
	delete first;
}

}  // namespace

extern "C" {

void ref_set_threads(int n) {
#ifdef _OPENMP
	omp_set_num_threads(n);
#endif
	(void)n;
}

int ref_max_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

// ChromosomeOneDigit(seq, header): Chromosome.cpp:40-44,99-112 + ChromosomeOneDigit.cpp:95-144.
// digits_out gets the encoded base string (len bytes); segs_out gets [start,end] pairs.
// returns the number of segments, or -1 if the reference throws.
int ref_encode(const char *seq, long len, char *digits_out, int *segs_out, int max_segs) {
	try {
		std::string s(seq, (size_t)len), info(">x");
		ChromosomeOneDigit chrom(s, info);
		const std::string *b = chrom.getBase();
		memcpy(digits_out, b->data(), b->size());
		auto *seg = chrom.getSegment();
		int n = (int)seg->size();
		for (int i = 0; i < n && i < max_segs; i++) {
			segs_out[2 * i] = seg->at(i)->at(0);
			segs_out[2 * i + 1] = seg->at(i)->at(1);
		}
		return n;
	} catch (...) {
		return -1;
	}
}

// fill_table<uint64_t> with KmerHashTable(k, 1): ClusterFactory.h:40-55, ClusterFactory.cpp:995,
// Runner.cpp:57-67.  out has 4^k entries.  returns 0 or -1 on throw.
int ref_hist_u64(const char *seq, long len, int k, uint64_t *out) {
	try {
		std::string s(seq, (size_t)len), info(">x");
		ChromosomeOneDigit chrom(s, info);
		KmerHashTable<unsigned long, uint64_t> table(k, 1);
		std::vector<uint64_t> values;
		fill_table<uint64_t>(table, &chrom, values);
		memcpy(out, values.data(), values.size() * sizeof(uint64_t));
		return 0;
	} catch (...) {
		return -1;
	}
}

// whole batches, threaded exactly like Runner.cpp:57 / ClusterFactory.cpp:785 (omp parallel for
// over sequences).  seqs = concatenated raw letters, offs[n+1].  out = n x 4^k of tbytes-wide bins.
int ref_hist_batch(const char *seqs, const int64_t *offs, int n, int k, int tbytes, void *out) {
	const size_t nb = (size_t)1 << (2 * k);
	int bad = 0;
#pragma omp parallel for schedule(dynamic, 16)
	for (int i = 0; i < n; i++) {
		try {
			std::string s(seqs + offs[i], (size_t)(offs[i + 1] - offs[i])), info(">x");
			ChromosomeOneDigit chrom(s, info);
			if (tbytes == 1) {
				KmerHashTable<unsigned long, uint8_t> table(k, 1);
				std::vector<uint8_t> v;
				fill_table<uint8_t>(table, &chrom, v);
				memcpy((uint8_t *)out + i * nb, v.data(), nb);
			} else if (tbytes == 2) {
				KmerHashTable<unsigned long, uint16_t> table(k, 1);
				std::vector<uint16_t> v;
				fill_table<uint16_t>(table, &chrom, v);
				memcpy((uint16_t *)out + i * nb, v.data(), nb * 2);
			} else {
				KmerHashTable<unsigned long, uint64_t> table(k, 1);
				std::vector<uint64_t> v;
				fill_table<uint64_t>(table, &chrom, v);
				memcpy((uint64_t *)out + i * nb, v.data(), nb * 8);
			}
		} catch (...) {
#pragma omp atomic write
			bad = 1;
		}
	}
	return bad ? -1 : 0;
}

// out5 = [LD, INTERSECTION, MANHATTAN, PEARSON, KULCZYNSKI2] raw (Feature.cpp), dist = distance()
void ref_features(const void *p, const void *q, int nbins, int tbytes, uint64_t lp, uint64_t lq,
		  double *out5, uint64_t *dist) {
	if (tbytes == 1) features_t<uint8_t>(p, q, nbins, lp, lq, out5, dist);
	else if (tbytes == 2) features_t<uint16_t>(p, q, nbins, lp, lq, out5, dist);
	else if (tbytes == 4) features_t<uint32_t>(p, q, nbins, lp, lq, out5, dist);
	else features_t<uint64_t>(p, q, nbins, lp, lq, out5, dist);
}

double ref_distance_d(const void *p, int nbins, int tbytes, const double *mean) {
	if (tbytes == 1) return distance_d_t<uint8_t>(p, nbins, mean);
	if (tbytes == 2) return distance_d_t<uint16_t>(p, nbins, mean);
	if (tbytes == 4) return distance_d_t<uint32_t>(p, nbins, mean);
	return distance_d_t<uint64_t>(p, nbins, mean);
}

void ref_mean(const void *hists, int nbins, int tbytes, int m, double *mean_out) {
	if (tbytes == 1) mean_t<uint8_t>(hists, nbins, m, mean_out);
	else if (tbytes == 2) mean_t<uint16_t>(hists, nbins, m, mean_out);
	else if (tbytes == 4) mean_t<uint32_t>(hists, nbins, m, mean_out);
	else mean_t<uint64_t>(hists, nbins, m, mean_out);
}

// A point-vs-center scan exactly as Trainer::get_close evaluates it (Trainer.cpp:81-106) with a
// Feature<T> whose bounds/combos were set like Trainer::train does (Trainer.cpp:584-587,606-611):
// combos f0=LD*INT, f1=(LD*MAN)^2, f2=PEARSON, f3=(LD*KUL)^2; lookup order [LD,INT,MAN,PEARSON,KUL]
// is what add_feature() produces (ascending bit order inside each call, Feature.cpp:15-28).  mins/maxs are given in that
// lookup order.  nfeat in {3,4}.  Outputs per point: sum, f0, flag.  threaded like the reference.
void ref_scan_u8(const uint8_t *hists, const uint64_t *lens, int n, int nbins,
		 const uint8_t *center, uint64_t center_len,
		 const double *mins, const double *maxs, const double *weights, int nfeat,
		 double *sum_out, double *f0_out, uint8_t *flag_out) {
	Feature<uint8_t> feat(0, NULL, 0);
	feat.add_feature(FEAT_INTERSECTION | FEAT_LD, COMBO_SELF);
	feat.add_feature(FEAT_MANHATTAN | FEAT_LD, COMBO_SQUARED);
	feat.add_feature(FEAT_PEARSON, COMBO_SELF);
	if (nfeat >= 4) feat.add_feature(FEAT_KULCZYNSKI2 | FEAT_LD, COMBO_SQUARED);
	for (size_t i = 0; i < feat.lookup.size(); i++) {
		feat.mins[i] = mins[i];
		feat.maxs[i] = maxs[i];
	}
	feat.finalize();
	DivergencePoint<uint8_t> *c = mk_point<uint8_t>(center, nbins, center_len);
	const int ncols = nfeat + 1;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++) {
		DivergencePoint<uint8_t> *pt = mk_point<uint8_t>(hists + (size_t)i * nbins, nbins, lens[i]);
		double sum = weights[0];
		double dist = 0;
		auto cache = feat.compute(*pt, *c);
		for (int col = 1; col < ncols; col++) {
			if (col == 1) {
				dist = feat(col - 1, cache);
				sum += weights[col] * dist;
			} else {
				sum += weights[col] * feat(col - 1, cache);
			}
		}
		double res = round(1.0 / (1 + exp(-sum)));
		sum_out[i] = sum;
		f0_out[i] = dist;
		flag_out[i] = (res == 1.0);
		delete pt;
	}
	delete c;
}

// The same evaluation over points that already exist as DivergencePoint<uint8_t> objects, the way
// the reference holds them (built once by ClusterFactory::build_points): this is the timed CPU
// baseline, so object construction stays outside it.
struct RefPointSet {
	std::vector<DivergencePoint<uint8_t> *> pts;
	Feature<uint8_t> *feat = nullptr;
	int nfeat = 0;
};

void *ref_pointset_create(const uint8_t *hists, const uint64_t *lens, int n, int nbins) {
	RefPointSet *ps = new RefPointSet();
	ps->pts.resize(n);
	for (int i = 0; i < n; i++) {
		ps->pts[i] = mk_point<uint8_t>(hists + (size_t)i * nbins, nbins, lens[i]);
		ps->pts[i]->set_id(i);
	}
	return ps;
}

void ref_pointset_destroy(void *h) {
	RefPointSet *ps = (RefPointSet *)h;
	for (auto *p : ps->pts) delete p;
	delete ps->feat;
	delete ps;
}

void ref_pointset_set_model(void *h, const double *mins, const double *maxs, int nfeat) {
	RefPointSet *ps = (RefPointSet *)h;
	delete ps->feat;
	ps->feat = new Feature<uint8_t>(0, NULL, 0);
	ps->feat->add_feature(FEAT_INTERSECTION | FEAT_LD, COMBO_SELF);
	ps->feat->add_feature(FEAT_MANHATTAN | FEAT_LD, COMBO_SQUARED);
	ps->feat->add_feature(FEAT_PEARSON, COMBO_SELF);
	if (nfeat >= 4) ps->feat->add_feature(FEAT_KULCZYNSKI2 | FEAT_LD, COMBO_SQUARED);
	for (size_t i = 0; i < ps->feat->lookup.size(); i++) {
		ps->feat->mins[i] = mins[i];
		ps->feat->maxs[i] = maxs[i];
	}
	ps->feat->finalize();
	ps->nfeat = nfeat;
}

// Trainer::get_close loop body + reductions (Trainer.cpp:81-106): returns the number of positives,
// *best_idx / *best_f0 the argmax of f0 (first max wins), flag_out optional.
long ref_pointset_scan(void *h, int center, const double *weights, uint8_t *flag_out, long *best_idx,
		       double *best_f0) {
	RefPointSet *ps = (RefPointSet *)h;
	Feature<uint8_t> &feat = *ps->feat;
	const int ncols = ps->nfeat + 1;
	const int n = (int)ps->pts.size();
	DivergencePoint<uint8_t> *c = ps->pts[center];
	long npos = 0;
	long bi = -1;
	double bf = -1;
#pragma omp parallel
	{
		long lpos = 0, lbi = -1;
		double lbf = -1;
#pragma omp for schedule(static) nowait
		for (int i = 0; i < n; i++) {
			DivergencePoint<uint8_t> *pt = ps->pts[i];
			double sum = weights[0];
			double dist = 0;
			auto cache = feat.compute(*pt, *c);
			for (int col = 1; col < ncols; col++) {
				if (col == 1) {
					dist = feat(col - 1, cache);
					sum += weights[col] * dist;
				} else {
					sum += weights[col] * feat(col - 1, cache);
				}
			}
			double res = round(1.0 / (1 + exp(-sum)));
			if (dist > lbf) { lbf = dist; lbi = i; }
			if (res == 1.0) lpos++;
			if (flag_out) flag_out[i] = (res == 1.0);
		}
#pragma omp critical
		{
			npos += lpos;
			if (lbi >= 0 && (lbf > bf || (lbf == bf && lbi < bi))) { bf = lbf; bi = lbi; }
		}
	}
	*best_idx = bi;
	*best_f0 = bf;
	return npos;
}

// GlobAlignE(seq1, 0, la-1, seq2, 0, lb-1, match, mismatch, open, cont): GlobAlignE.cpp:22-57,123-305
void ref_globalign(const char *s1, int la, const char *s2, int lb,
		   int match, int mismatch, int gopen, int gcont,
		   int *score, int *alen, int *matches, double *identity) {
	utility::GlobAlignE g(s1, 0, la - 1, s2, 0, lb - 1, match, mismatch, gopen, gcont);
	*score = g.getScore();
	*alen = g.getLength();
	*matches = g.totalMatches;
	*identity = g.getIdentity();
}

// batch of alignments over digit strings (Trainer::align / Feature::align parameters 1,-1,2,1),
// threaded like Trainer::get_labels (Trainer.cpp:282).
void ref_globalign_batch(const char *seqs, const int64_t *offs, const int32_t *pa, const int32_t *pb,
			 int npairs, int *score, int *alen, int *matches) {
#pragma omp parallel for schedule(dynamic)
	for (int i = 0; i < npairs; i++) {
		int a = pa[i], b = pb[i];
		int la = (int)(offs[a + 1] - offs[a]), lb = (int)(offs[b + 1] - offs[b]);
		utility::GlobAlignE g(seqs + offs[a], 0, la - 1, seqs + offs[b], 0, lb - 1, 1, -1, 2, 1);
		score[i] = g.getScore();
		alen[i] = g.getLength();
		matches[i] = g.totalMatches;
	}
}

}  // extern "C"
