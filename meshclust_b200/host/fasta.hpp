// Multi-FASTA reader with the record semantics of ChromListMaker::makeChromOneDigitList
// (nonltr/ChromListMaker.cpp:92-120): a line starting with '>' opens a record and is kept whole
// (including '>' and the description) as the header; every other line is appended verbatim to the
// sequence; lines end at \n, \r\n or \r (safe_getline, :23-47).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace mch {

struct FastaBatch {
	std::vector<std::string> headers;
	std::vector<uint8_t> letters;     // concatenated raw sequence bytes (no newlines)
	std::vector<int64_t> offsets{0};  // n+1
	size_t size() const { return headers.size(); }
};

// appends the records of `path` to `out`; returns false (with msg) when the file cannot be used
inline bool read_fasta(const std::string &path, FastaBatch &out, std::string &msg) {
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) { msg = "File \"" + path + "\" does not exist"; return false; }
	std::vector<char> buf;
	{
		fseek(f, 0, SEEK_END);
		const long sz = ftell(f);
		fseek(f, 0, SEEK_SET);
		buf.resize(sz > 0 ? (size_t)sz : 0);
		if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); msg = "short read on " + path; return false; }
		fclose(f);
	}
	const size_t n = buf.size();
	size_t i = 0;
	bool open = false, has_line = false;
	auto close_record = [&]() -> bool {
		if (!open) return true;
		if (!has_line) {   // Chromosome::finalize: header and sequence must both have been set
			msg = "record \"" + out.headers.back() + "\" has no sequence line";
			return false;
		}
		out.offsets.push_back((int64_t)out.letters.size());
		return true;
	};
	// the reference loops `while (in.good())`: after the last newline it reads one more, empty, line
	bool more = true;
	while (more) {
		size_t j = i;
		while (j < n && buf[j] != '\n' && buf[j] != '\r') j++;
		const char *line = buf.data() + i;
		const size_t len = j - i;
		if (j >= n) more = false;                       // EOF reached while reading this line
		else if (buf[j] == '\r' && j + 1 < n && buf[j + 1] == '\n') i = j + 2;
		else i = j + 1;
		if (len > 0 && line[0] == '>') {
			if (!close_record()) return false;
			out.headers.emplace_back(line, len);
			open = true;
			has_line = false;
		} else {
			if (!open) {
				if (len == 0 && !more) break;            // empty file
				msg = "sequence data before the first header in " + path;
				return false;
			}
			out.letters.insert(out.letters.end(), line, line + len);
			has_line = true;
		}
	}
	if (!open) { msg = "no FASTA record in " + path; return false; }
	return close_record();
}

}  // namespace mch
