// Dense double matrices for the <= 5x5 "GLM" solve (host side, not a kernel).
// Follows the arithmetic ORDER of the reference's matrix::Matrix / matrix::GLM
// (Matrix.cpp:69-89 multiply i,j,k; :102-200 Gauss-Jordan; :202-214 pseudo-inverse; GLM.cpp:19-34)
// so the trained weights agree to the last bits: the reference binary evaluates the multiply
// with separate mul/add and the elimination step x - p*y as one fused negate-multiply-add
// (objdump of the -O3 build), which is what the explicit std::fma below reproduces; this file is
// compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdio>
#include <vector>

namespace mch {

struct Mat {
	int r = 0, c = 0;
	std::vector<double> v;
	Mat() {}
	Mat(int rows, int cols) : r(rows), c(cols), v((size_t)rows * cols, 0.0) {}
	double &at(int i, int j) { return v[(size_t)i * c + j]; }
	double at(int i, int j) const { return v[(size_t)i * c + j]; }
};

inline Mat mul(const Mat &a, const Mat &b) {
	Mat out(a.r, b.c);
	for (int i = 0; i < out.r; i++)
		for (int j = 0; j < out.c; j++) {
			double s = 0;
			for (int k = 0; k < a.c; k++) {
				const double prod = a.at(i, k) * b.at(k, j);
				s = s + prod;
			}
			out.at(i, j) = s;
		}
	return out;
}

inline Mat transpose(const Mat &a) {
	Mat t(a.c, a.r);
	for (int i = 0; i < a.r; i++)
		for (int j = 0; j < a.c; j++) t.at(j, i) = a.at(i, j);
	return t;
}

// Gauss-Jordan with the reference's pivoting rules: scale the pivot row unless the pivot is
// exactly 1; on an exactly-zero pivot swap with the first lower row whose entry is non-zero;
// eliminate below, then above from the last column up; finally demand an exact identity, else
// "Inverse does not exist" and the ORIGINAL matrix is returned (Matrix.cpp:114-115,184-196).
inline Mat gauss_jordan_inverse(const Mat &orig) {
	const int n = orig.r;
	Mat a = orig, inv(n, n);
	for (int i = 0; i < n; i++) inv.at(i, i) = 1;
	auto fail = [&]() {
		printf("Inverse does not exist\n");
		return orig;
	};
	for (int i = 0; i < n; i++) {
		if (a.at(i, i) != 1) {
			if (a.at(i, i) == 0) {
				int row = i + 1;
				while (row < n && a.at(row, i) == 0) row++;
				if (row >= n) return fail();
				for (int j = 0; j < n; j++) {
					std::swap(a.at(i, j), a.at(row, j));
					std::swap(inv.at(i, j), inv.at(row, j));
				}
			}
			const double p = a.at(i, i);
			for (int j = 0; j < n; j++) {
				a.at(i, j) = a.at(i, j) / p;
				inv.at(i, j) = inv.at(i, j) / p;
			}
		}
		for (int below = i + 1; below < n; below++) {
			if (a.at(below, i) != 0) {
				const double p = a.at(below, i);
				for (int j = 0; j < n; j++) {
					a.at(below, j) = std::fma(-p, a.at(i, j), a.at(below, j));
					inv.at(below, j) = std::fma(-p, inv.at(i, j), inv.at(below, j));
				}
			}
		}
	}
	for (int i = n - 1; i >= 0; i--) {
		for (int above = 0; above < i; above++) {
			if (a.at(above, i) != 0) {
				const double p = a.at(above, i);
				for (int j = 0; j < n; j++) {
					a.at(above, j) = std::fma(-p, a.at(i, j), a.at(above, j));
					inv.at(above, j) = std::fma(-p, inv.at(i, j), inv.at(above, j));
				}
			}
		}
	}
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++)
			if ((i == j && a.at(i, j) != 1) || (i != j && a.at(i, j) != 0)) return fail();
	return inv;
}

// rows >= cols branch of Matrix::pseudoInverse: (A^T A)^-1 A^T
inline Mat pseudo_inverse(const Mat &a) {
	const Mat t = transpose(a);
	if (a.r >= a.c) return mul(gauss_jordan_inverse(mul(t, a)), t);
	return mul(t, gauss_jordan_inverse(mul(a, t)));
}

// GLM::train: w = pinv(X^T X) * X^T * y   (left to right)
inline Mat glm_train(const Mat &X, const Mat &y) {
	const Mat xt = transpose(X);
	const Mat xtx = mul(xt, X);
	return mul(mul(pseudo_inverse(xtx), xt), y);
}

// GLM::predict + the 0 -> -1 relabel of Trainer::train; returns accuracy in percent
inline double glm_accuracy(const Mat &X, const Mat &w, const Mat &labels, bool print = true) {
	const Mat p = mul(X, w);
	int sum = 0, neg = 0, negsame = 0, pos = 0, possame = 0;
	for (int i = 0; i < X.r; i++) {
		double lab = std::round(1 / (1 + std::exp(-p.at(i, 0))));
		if (lab == 0) lab = -1;
		if (labels.at(i, 0) == -1) {
			neg++;
			if (lab == -1) { sum++; negsame++; }
		} else {
			pos++;
			if (labels.at(i, 0) == lab) { sum++; possame++; }
		}
	}
	const double acc = ((double)sum * 100) / X.r;
	if (print)
		printf("Accuracy: %g%% Sensitivity: %g%% Specificity: %g%% \n", acc, ((double)possame * 100) / pos,
		       ((double)negsame * 100) / neg);
	return acc;
}

}  // namespace mch
