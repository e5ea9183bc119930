// Self-check of host/fasta.hpp: the parallel parser must produce exactly what the serial parser
// does on well-formed files, and must hand every irregular file back to the serial parser.  The indexer
// (what the device-side ingest is fed with) must describe the same records: squeezing the line feeds out
// of every span gives the serial parser's letters, and the announced letter counts are exact.
#include <cstdio>
#include <random>
#include <string>

#include "../../meshclust_b200/host/fasta.hpp"

static bool same(const mch::FastaBatch &a, const mch::FastaBatch &b) {
	if (a.headers != b.headers || a.offsets != b.offsets || a.letters.size() != b.letters.size()) return false;
	return memcmp(a.letters.data(), b.letters.data(), a.letters.size()) == 0;
}

int main(int argc, char **argv) {
	const std::string dir = argc > 1 ? argv[1] : "/tmp";
	std::mt19937 rng(7);
	int fast = 0;
	for (int round = 0; round < 60; round++) {
		std::string txt;
		const int nrec = 1 + rng() % 400;
		const int kind = round % 6;   // 0-2 regular, 3 CRLF, 4 header without sequence, 5 no trailing newline
		for (int r = 0; r < nrec; r++) {
			txt += ">seq" + std::to_string(r) + " x" + std::to_string(rng() % 1000) + (kind == 3 ? "\r\n" : "\n");
			if (kind == 4 && r == nrec / 2) continue;
			const int len = rng() % 500;
			for (int i = 0; i < len; i++) {
				txt += "ACGTNacgtRY"[rng() % 11];
				if (i % 70 == 69) txt += kind == 3 ? "\r\n" : "\n";
			}
			if (rng() % 5 == 0) txt += "\n";   // empty line inside a record
			txt += kind == 3 ? "\r\n" : "\n";
		}
		if (kind == 5) while (!txt.empty() && txt.back() == '\n') txt.pop_back();
		const std::string path = dir + "/fasta_selfcheck.fa";
		FILE *f = fopen(path.c_str(), "wb");
		fwrite(txt.data(), 1, txt.size(), f);
		fclose(f);
		mch::FastaBatch a, b;
		std::string ma, mb;
		const bool oa = mch::read_fasta(path, a, ma, (size_t)1 << 60);   // serial only
		const bool ob = mch::read_fasta(path, b, mb, 1);                // parallel whenever it applies
		if (oa != ob || ma != mb || (oa && !same(a, b))) { printf("mismatch in round %d (kind %d): %d/%d '%s' '%s'\n", round, kind, oa, ob, ma.c_str(), mb.c_str()); return 1; }
		mch::FastaBatch c;
		if (mch::parse_fasta_parallel(txt.data(), txt.size(), c, 5)) fast++;
		// appending a second file must also agree
		if (oa) {
			const bool oa2 = mch::read_fasta(path, a, ma, (size_t)1 << 60), ob2 = mch::read_fasta(path, b, mb, 1);
			if (!oa2 || !ob2 || !same(a, b)) { printf("append mismatch in round %d\n", round); return 1; }
		}
	}
	// the indexer: two files per round (the second one a copy with other record names), against the serial parser
	int indexed = 0;
	for (int round = 0; round < 60; round++) {
		std::string txt[2];
		const int kind = round % 6;
		for (int fi = 0; fi < 2; fi++) {
			const int nrec = 1 + rng() % (round < 6 ? 3 : 300);
			for (int r = 0; r < nrec; r++) {
				txt[fi] += ">f" + std::to_string(fi) + "r" + std::to_string(r) + (kind == 3 ? "\r\n" : "\n");
				if (kind == 4 && fi == 1 && r == nrec / 2) continue;
				const int len = rng() % 300, width = 1 + rng() % 90;
				for (int i = 0; i < len; i++) {
					txt[fi] += "ACGTNacgtRY>"[rng() % (i % width == 0 ? 11 : 12)];   // a '>' inside a line is a letter
					if (i % width == width - 1) txt[fi] += kind == 3 ? "\r\n" : "\n";
				}
				if (rng() % 5 == 0) txt[fi] += "\n";
				txt[fi] += kind == 3 ? "\r\n" : "\n";
			}
			if (kind == 5) while (!txt[fi].empty() && txt[fi].back() == '\n') txt[fi].pop_back();
		}
		std::vector<std::string> paths;
		for (int fi = 0; fi < 2; fi++) {
			paths.push_back(dir + "/fasta_selfcheck_" + std::to_string(fi) + ".fa");
			FILE *f = fopen(paths.back().c_str(), "wb");
			fwrite(txt[fi].data(), 1, txt[fi].size(), f);
			fclose(f);
		}
		mch::FastaBatch a;
		std::string msg;
		bool oa = true;
		std::vector<size_t> first;
		for (int fi = 0; fi < 2 && oa; fi++) { first.push_back(a.size()); oa = mch::read_fasta(paths[fi], a, msg, (size_t)1 << 60); }
		first.push_back(a.size());
		mch::FastaIndex ix;
		const bool oi = mch::index_fasta_files(paths, ix);
		if (oi && !oa) { printf("the indexer accepted what the parser rejects (round %d, kind %d)\n", round, kind); return 1; }
		if (!oi) {
			if (kind != 3 && kind != 4) { printf("the indexer refused a regular input (round %d, kind %d)\n", round, kind); return 1; }
			continue;
		}
		indexed++;
		if (ix.headers != a.headers || ix.file_first != first || ix.raw.size() != txt[0].size() + txt[1].size()) { printf("index: headers / files differ in round %d\n", round); return 1; }
		for (size_t r = 0; r < ix.size(); r++) {
			std::string sq;
			for (int64_t p = ix.span_begin[r]; p < ix.span_end[r]; p++) if (ix.raw.data()[p] != '\n') sq += (char)ix.raw.data()[p];
			const std::string want((const char *)a.letters.data() + a.offsets[r], (size_t)(a.offsets[r + 1] - a.offsets[r]));
			if (sq != want || (int64_t)sq.size() != ix.letters[r]) { printf("index: record %zu differs in round %d (kind %d)\n", r, round, kind); return 1; }
		}
	}
	if (indexed < 30) { printf("indexer used only %d times\n", indexed); return 1; }
	if (fast < 20) { printf("parallel path taken only %d times\n", fast); return 1; }
	printf("ok (%d files took the parallel path)\n", fast);
	return 0;
}
