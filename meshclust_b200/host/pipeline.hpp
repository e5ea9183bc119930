// bin/meshclust host side: the reference's Runner / Trainer / ClusterFactory control flow
// re-created on top of the meshclust_b200 C-ABI (every arithmetic-heavy loop is a GPU call).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mch {

// Runner.h:23-33 + Runner.cpp:28-38 defaults
struct Options {
	int k = -1;
	double similarity = 0.90;
	int iterations = 15;
	int delta = 5;
	bool align = false;
	int sample_size = 0;   // 0 -> 3000
	int pivots = 20;
	int threads = 0;
	int device = 0;
	int gpus = 1;          // extension: GPUs that share the Phase-A scans
	std::vector<std::string> files;
	std::string output = "output.clstr";
	std::string dump_model;   // test hook: write the trained bounds/weights/sample sizes as text
};

// parses argv exactly like Runner::get_opts (Runner.cpp:150-263); exits like the reference on errors
Options parse_options(int argc, char **argv);
void usage(const std::string &prog);

int run_pipeline(Options opt);

}  // namespace mch
