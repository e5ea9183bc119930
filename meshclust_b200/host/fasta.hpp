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

}  // namespace mch
