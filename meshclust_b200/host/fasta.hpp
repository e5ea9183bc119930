// Multi-FASTA reader with the record semantics of ChromListMaker::makeChromOneDigitList
// (nonltr/ChromListMaker.cpp:92-120): a line starting with '>' opens a record and is kept whole
// (including '>' and the description) as the header; every other line is appended verbatim to the
// sequence; lines end at \n, \r\n or \r (safe_getline, :23-47).
#pragma once
#include <cstdint>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace mch {

// byte buffer that grows without value-initialising (a 1 GB std::vector::resize is a 1 GB memset)
class RawBytes {
public:
	RawBytes() = default;
	RawBytes(const RawBytes &) = delete;
	RawBytes &operator=(const RawBytes &) = delete;
	~RawBytes() { free(p_); }
	uint8_t *data() { return p_; }
	const uint8_t *data() const { return p_; }
	size_t size() const { return n_; }
	void release() { free(p_); p_ = nullptr; n_ = cap_ = 0; }
	void resize(size_t n) {
		if (n > cap_) {
			void *q = realloc(p_, n);
			if (!q) throw std::bad_alloc();
			p_ = (uint8_t *)q;
			cap_ = n;
		}
		n_ = n;
	}
private:
	uint8_t *p_ = nullptr;
	size_t n_ = 0, cap_ = 0;
};

struct FastaBatch {
	std::vector<std::string> headers;
	RawBytes letters;                 // concatenated raw sequence bytes (no newlines)
	std::vector<int64_t> offsets{0};  // n+1
	size_t size() const { return headers.size(); }
};

// Parallel fast path for well-formed LF-only files: the buffer is cut into one piece per thread, a
// line belongs to the piece that holds its first byte; pass 1 counts headers and letters per piece,
// a prefix sum places every piece, pass 2 copies.  Returns false when the file needs the serial
// parser (CR line ends, data before the first header, a header without a sequence line, ...), which
// is the single source of truth for the error semantics.
inline bool parse_fasta_parallel(const char *bd, size_t n, FastaBatch &out, int pieces) {
	if (n == 0 || bd[0] != '>' || pieces < 2) return false;
	struct Piece { size_t begin = 0, end = 0, headers = 0, letters = 0; bool bad = false; };
	std::vector<Piece> pc((size_t)pieces);
	bool has_cr = false;
#pragma omp parallel for schedule(static) reduction(|| : has_cr)
	for (int t = 0; t < pieces; t++) {
		size_t a = n * (size_t)t / (size_t)pieces, b = n * (size_t)(t + 1) / (size_t)pieces;
		if (memchr(bd + a, '\r', b - a)) has_cr = true;
		// first line start at or after a
		if (t > 0) {
			const char *nl = (const char *)memchr(bd + a - 1, '\n', n - (a - 1));
			a = nl ? (size_t)(nl - bd) + 1 : n;
		}
		if (t + 1 < pieces) {
			const char *nl = b > 0 ? (const char *)memchr(bd + b - 1, '\n', n - (b - 1)) : nullptr;
			b = nl ? (size_t)(nl - bd) + 1 : n;
		}
		Piece &p = pc[(size_t)t];
		p.begin = a; p.end = b;
		for (size_t i = a; i < b;) {
			const char *nl = (const char *)memchr(bd + i, '\n', n - i);
			const size_t j = nl ? (size_t)(nl - bd) : n;
			if (bd[i] == '>' && j > i) {
				p.headers++;
				// a header needs a following line that is not a header (an empty one counts: safe_getline)
				if (j >= n || (j + 1 < n && bd[j + 1] == '>')) p.bad = true;
			} else p.letters += j - i;
			i = j + 1;
		}
	}
	if (has_cr) return false;
	size_t nh = 0, nl = 0;
	for (auto &p : pc) { if (p.bad) return false; nh += p.headers; nl += p.letters; }
	const size_t rec0 = out.headers.size(), base = out.letters.size();
	out.headers.resize(rec0 + nh);
	out.offsets.resize(rec0 + nh + 1);
	out.letters.resize(base + nl);
	std::vector<size_t> hoff((size_t)pieces + 1, 0), loff((size_t)pieces + 1, 0);
	for (int t = 0; t < pieces; t++) { hoff[t + 1] = hoff[t] + pc[t].headers; loff[t + 1] = loff[t] + pc[t].letters; }
	uint8_t *dst = out.letters.data();
#pragma omp parallel for schedule(static)
	for (int t = 0; t < pieces; t++) {
		const Piece &p = pc[(size_t)t];
		size_t rec = rec0 + hoff[t], w = base + loff[t];
		for (size_t i = p.begin; i < p.end;) {
			const char *e = (const char *)memchr(bd + i, '\n', n - i);
			const size_t j = e ? (size_t)(e - bd) : n;
			if (bd[i] == '>' && j > i) {
				out.headers[rec].assign(bd + i, j - i);
				out.offsets[rec] = (int64_t)w;   // the record's letters start here
				rec++;
			} else {
				memcpy(dst + w, bd + i, j - i);
				w += j - i;
			}
			i = j + 1;
		}
	}
	out.offsets[rec0 + nh] = (int64_t)(base + nl);
	return true;
}

// appends the records of `path` to `out`; returns false (with msg) when the file cannot be used
inline bool read_fasta(const std::string &path, FastaBatch &out, std::string &msg, size_t parallel_min_bytes = (size_t)32 << 20) {
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) { msg = "File \"" + path + "\" does not exist"; return false; }
	RawBytes buf;
	{
		fseek(f, 0, SEEK_END);
		const long sz = ftell(f);
		fseek(f, 0, SEEK_SET);
		buf.resize(sz > 0 ? (size_t)sz : 0);
		bool short_read = false;
		if (sz > 0 && (size_t)sz >= parallel_min_bytes) {
			// page-cache copies are memcpy-bound: let the host threads share them
			const int fd = fileno(f);
			const int parts = 16;
#pragma omp parallel for schedule(dynamic) reduction(|| : short_read)
			for (int t = 0; t < parts; t++) {
				size_t a = (size_t)sz * (size_t)t / parts;
				const size_t b = (size_t)sz * (size_t)(t + 1) / parts;
				while (a < b) {
					const ssize_t got = pread(fd, buf.data() + a, b - a, (off_t)a);
					if (got <= 0) { short_read = true; break; }
					a += (size_t)got;
				}
			}
		} else if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) short_read = true;
		fclose(f);
		if (short_read) { msg = "short read on " + path; return false; }
	}
	const size_t n = buf.size();
	const char *bd = (const char *)buf.data();
	if (n >= parallel_min_bytes) {
		int pieces = 1;
#ifdef _OPENMP
		pieces = omp_get_max_threads();
#endif
		if (parse_fasta_parallel(bd, n, out, pieces)) return true;
	}
	size_t i = 0;
	bool open = false, has_line = false;
	const size_t base = out.letters.size();
	size_t w = base;
	auto close_record = [&]() -> bool {
		if (!open) return true;
		if (!has_line) {   // Chromosome::finalize: header and sequence must both have been set
			msg = "record \"" + out.headers.back() + "\" has no sequence line";
			return false;
		}
		out.offsets.push_back((int64_t)w);
		return true;
	};
	// sequence bytes are compacted into one buffer sized for the whole file up front (the letters
	// of a file never outnumber its bytes); line ends are found with memchr
	out.letters.resize(base + n);
	uint8_t *dst = out.letters.data();
	const bool any_cr = n > 0 && memchr(bd, '\r', n) != nullptr;
	// the reference loops `while (in.good())`: after the last newline it reads one more, empty, line
	bool more = true;
	while (more) {
		size_t j;
		{
			const char *nl = i < n ? (const char *)memchr(bd + i, '\n', n - i) : nullptr;
			j = nl ? (size_t)(nl - bd) : n;
			if (any_cr && j > i) {
				const char *cr = (const char *)memchr(bd + i, '\r', j - i);
				if (cr) j = (size_t)(cr - bd);
			}
		}
		const char *line = bd + i;
		const size_t len = j - i;
		if (j >= n) more = false;                       // EOF reached while reading this line
		else if (bd[j] == '\r' && j + 1 < n && bd[j + 1] == '\n') i = j + 2;
		else i = j + 1;
		if (len > 0 && line[0] == '>') {
			if (!close_record()) return false;
			out.headers.emplace_back(line, len);
			open = true;
			has_line = false;
		} else {
			if (!open) {
				if (len == 0 && !more) break;            // empty file
				out.letters.resize(base);
				msg = "sequence data before the first header in " + path;
				return false;
			}
			memcpy(dst + w, line, len);
			w += len;
			has_line = true;
		}
	}
	out.letters.resize(w);
	if (!open) { msg = "no FASTA record in " + path; return false; }
	return close_record();
}

// ---- index only: what the device-side ingest (mc_ingest_fasta) needs --------------------------------------------
// The files' bytes stay as they are (one buffer, file after file); per record the header, the byte span of its
// sequence lines and the number of letters in it.  One parallel pass over the bytes, no copy of the letters.
struct FastaIndex {
	RawBytes raw;                        // all files, concatenated
	std::vector<std::string> headers;
	std::vector<int64_t> span_begin, span_end;   // sequence lines of record i: raw[span_begin[i], span_end[i])
	std::vector<int64_t> letters;        // letters in that span (bytes that are not '\n')
	std::vector<size_t> file_first;      // first record of every file, then the record count
	size_t size() const { return headers.size(); }
	void clear() {
		raw.release();
		std::vector<std::string>().swap(headers);
		std::vector<int64_t>().swap(span_begin);
		std::vector<int64_t>().swap(span_end);
		std::vector<int64_t>().swap(letters);
		file_first.clear();
	}
};

// One file that already sits at raw[base, base + n).  Same eligibility as parse_fasta_parallel: LF line ends, the
// file starts with a header, every header is followed by a line that is not a header; false = use read_fasta.
inline bool index_fasta_region(const char *all, size_t base, size_t n, FastaIndex &out, int pieces) {
	const char *bd = all + base;
	if (n == 0 || bd[0] != '>') return false;
	if (pieces < 1) pieces = 1;
	if ((size_t)pieces > n / 4096 + 1) pieces = (int)(n / 4096 + 1);
	struct Piece {
		std::vector<size_t> hpos, hend, letters;   // per header that starts in the piece: '>' position, end of its line, letters after it (inside the piece)
		size_t lead = 0;                           // letters before the piece's first header: they belong to the record before
		bool bad = false, cr = false;
	};
	std::vector<Piece> pc((size_t)pieces);
#pragma omp parallel for schedule(static)
	for (int t = 0; t < pieces; t++) {
		size_t a = n * (size_t)t / (size_t)pieces, b = n * (size_t)(t + 1) / (size_t)pieces;
		Piece &p = pc[(size_t)t];
		if (b > a && memchr(bd + a, '\r', b - a)) p.cr = true;
		// a line belongs to the piece that holds its first byte
		if (t > 0) {
			const char *nl = (const char *)memchr(bd + a - 1, '\n', n - (a - 1));
			a = nl ? (size_t)(nl - bd) + 1 : n;
		}
		if (t + 1 < pieces) {
			const char *nl = b > 0 ? (const char *)memchr(bd + b - 1, '\n', n - (b - 1)) : nullptr;
			b = nl ? (size_t)(nl - bd) + 1 : n;
		}
		// i is always a line start.  A header line is taken as a line; the sequence lines behind it are taken as one
		// block up to the next line that starts with '>' (or the end of the piece): its letters are its bytes minus its
		// line feeds -- two vectorised scans per record instead of one memchr per 70-letter line
		for (size_t i = a; i < b;) {
			if (bd[i] == '>') {
				const char *nl = (const char *)memchr(bd + i, '\n', n - i);
				const size_t j = nl ? (size_t)(nl - bd) : n;
				p.hpos.push_back(i);
				p.hend.push_back(j);
				p.letters.push_back(0);
				// a header needs a following line that is not a header (an empty one counts: safe_getline)
				if (j >= n || (j + 1 < n && bd[j + 1] == '>')) p.bad = true;
				i = j + 1;
				continue;
			}
			size_t end = b;
			for (size_t q = i;;) {   // the next '>' that opens a line ('>' inside a line is a letter)
				const char *g = q < b ? (const char *)memchr(bd + q, '>', b - q) : nullptr;
				if (!g) break;
				const size_t gp = (size_t)(g - bd);
				if (bd[gp - 1] == '\n') { end = gp; break; }   // (gp > i >= a: bd[gp - 1] exists; i itself is not '>')
				q = gp + 1;
			}
			size_t feeds = 0;
			{
				const char *pb = bd + i;
				const size_t len = end - i;
				// (blocks of 4 KB so that the per-byte counts fit the vector lanes' 8-bit partial sums the compiler uses)
				for (size_t o = 0; o < len; o += 4096) {
					const size_t m = len - o < 4096 ? len - o : 4096;
					unsigned c = 0;
#pragma omp simd reduction(+ : c)
					for (size_t k2 = 0; k2 < m; k2++) c += pb[o + k2] == '\n' ? 1u : 0u;
					feeds += c;
				}
			}
			const size_t letters = (end - i) - feeds;
			if (p.letters.empty()) p.lead += letters;
			else p.letters.back() += letters;
			i = end;
		}
	}
	size_t nh = 0;
	for (const Piece &p : pc) {
		if (p.bad || p.cr) return false;
		nh += p.hpos.size();
	}
	const size_t rec0 = out.headers.size();
	out.headers.resize(rec0 + nh);
	out.span_begin.resize(rec0 + nh);
	out.span_end.resize(rec0 + nh);
	out.letters.resize(rec0 + nh);
	std::vector<size_t> first((size_t)pieces + 1, 0);
	for (int t = 0; t < pieces; t++) first[(size_t)t + 1] = first[(size_t)t] + pc[(size_t)t].hpos.size();
#pragma omp parallel for schedule(static)
	for (int t = 0; t < pieces; t++) {
		const Piece &p = pc[(size_t)t];
		for (size_t h = 0; h < p.hpos.size(); h++) {
			const size_t rec = rec0 + first[(size_t)t] + h;
			out.headers[rec].assign(bd + p.hpos[h], p.hend[h] - p.hpos[h]);
			out.span_begin[rec] = (int64_t)(base + std::min(p.hend[h] + 1, n));
			out.letters[rec] = (int64_t)p.letters[h];
			if (rec > rec0) out.span_end[rec - 1] = (int64_t)(base + p.hpos[h]);
		}
	}
	out.span_end[rec0 + nh - 1] = (int64_t)(base + n);
	// letters in front of a piece's first header belong to the last record opened before the piece
	size_t last_rec = rec0;
	for (int t = 0; t < pieces; t++) {
		const Piece &p = pc[(size_t)t];
		if (t > 0) out.letters[last_rec] += (int64_t)p.lead;   // (piece 0 starts with a header: lead = 0)
		if (!p.hpos.empty()) last_rec = rec0 + first[(size_t)t + 1] - 1;
	}
	return true;
}

// all files into one buffer + their index; false (nothing usable in `out`) when a file needs the serial parser
inline bool index_fasta_files(const std::vector<std::string> &files, FastaIndex &out) {
	std::vector<size_t> sizes;
	size_t total = 0;
	for (const std::string &path : files) {
		FILE *f = fopen(path.c_str(), "rb");
		if (!f) return false;
		fseek(f, 0, SEEK_END);
		const long sz = ftell(f);
		fclose(f);
		if (sz <= 0) return false;
		sizes.push_back((size_t)sz);
		total += (size_t)sz;
	}
	out.raw.resize(total);
	int threads = 1;
#ifdef _OPENMP
	threads = omp_get_max_threads();
#endif
	size_t base = 0;
	for (size_t fi = 0; fi < files.size(); fi++) {
		FILE *f = fopen(files[fi].c_str(), "rb");
		if (!f) return false;
		const int fd = fileno(f);
		const size_t sz = sizes[fi];
		bool short_read = false;
		const int parts = sz >= ((size_t)8 << 20) ? 16 : 1;   // page-cache copies are memcpy-bound: the host threads share them
#pragma omp parallel for schedule(dynamic) reduction(|| : short_read)
		for (int t = 0; t < parts; t++) {
			size_t a = sz * (size_t)t / (size_t)parts;
			const size_t b = sz * (size_t)(t + 1) / (size_t)parts;
			while (a < b) {
				const ssize_t got = pread(fd, out.raw.data() + base + a, b - a, (off_t)a);
				if (got <= 0) { short_read = true; break; }
				a += (size_t)got;
			}
		}
		fclose(f);
		if (short_read) return false;
		out.file_first.push_back(out.size());
		if (!index_fasta_region((const char *)out.raw.data(), base, sz, out, threads)) return false;
		base += sz;
	}
	out.file_first.push_back(out.size());
	return true;
}

}  // namespace mch
