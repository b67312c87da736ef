// TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
// Boost-free replacement for the reference's src/AlignerMain.cpp: fills
// AlignerParams exactly as AlignerMain.cpp:138-262 does for the options the
// alignment path uses and calls the reference's own alignReads() (Aligner.cpp:1124).
// Everything below main() is the UNMODIFIED reference, compiled from
// /root/reference by oracle/Makefile.
#include <csignal>
#include <omp.h>
#include "ref_params.h"
#include "ThreadReadAssertion.h"

int main(int argc, char** argv)
{
	struct sigaction act;
	act.sa_handler = ThreadReadAssertion::signal;
	sigemptyset(&act.sa_mask);
	act.sa_flags = 0;
	sigaction(SIGSEGV, &act, 0);
	std::vector<std::string> extra;
	AlignerParams params = gcParseArgs(argc, argv, extra);
	if (params.graphFile == "" || params.fastqFiles.empty() || (params.outputGAMFile == "" && params.outputJSONFile == "" && params.outputGAFFile == ""))
	{
		std::cerr << "graph file, read file and alignments-out must be given" << std::endl << "run with option -h for help" << std::endl;
		return 1;
	}
	omp_set_num_threads(params.numThreads);
	alignReads(params);
	return 0;
}
