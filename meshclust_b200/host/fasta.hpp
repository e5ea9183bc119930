// Multi-FASTA reader with the record semantics of ChromListMaker::makeChromOneDigitList
// (nonltr/ChromListMaker.cpp:92-120): a line starting with '>' opens a record and is kept whole
// (including '>' and the description) as the header; every other line is appended verbatim to the
// sequence; lines end at \n, \r\n or \r (safe_getline, :23-47).
#pragma once
#include <cstdint>
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

// appends the records of `path` to `out`; returns false (with msg) when the file cannot be used
inline bool read_fasta(const std::string &path, FastaBatch &out, std::string &msg) {
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) { msg = "File \"" + path + "\" does not exist"; return false; }
	RawBytes buf;
	{
		fseek(f, 0, SEEK_END);
		const long sz = ftell(f);
		fseek(f, 0, SEEK_SET);
		buf.resize(sz > 0 ? (size_t)sz : 0);
		if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); msg = "short read on " + path; return false; }
		fclose(f);
	}
	const size_t n = buf.size();
	const char *bd = (const char *)buf.data();
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
