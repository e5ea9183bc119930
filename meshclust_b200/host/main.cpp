// bin/meshclust -- drop-in for the reference CLI (src/cluster/src/main.cpp:8-27): same flags,
// same CD-HIT CLSTR output; the hot path runs on the GPU through the meshclust_b200 C-ABI.
#include "pipeline.hpp"

int main(int argc, char **argv) {
	mch::Options opt = mch::parse_options(argc, argv);
	return mch::run_pipeline(opt);
}
