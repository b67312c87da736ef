// gc_buildindex graph.gfa out.gcidx -- build and store the graph / MPC / minimizer index
// (host only; the on-disk index cache the reference left as a stub, AlignmentGraph.cpp:1490-1495).
#include <iostream>
#include "gc_builder.h"

int main(int argc, char** argv)
{
	if (argc < 3) { std::cerr << "usage: gc_buildindex graph.gfa out.gcidx [--quiet]" << std::endl; return 1; }
	bool verbose = !(argc > 3 && std::string(argv[3]) == "--quiet");
	try
	{
		GcIndexFile idx = gcbuild::buildIndexFromGfa(argv[1], 15, 20, 0.001, verbose);
		idx.save(argv[2]);
	}
	catch (const std::exception& e) { std::cerr << "Error in the graph: " << e.what() << std::endl; return 1; }
	return 0;
}
