// TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
// Boost-free restatement of the option handling of the reference's
// src/AlignerMain.cpp:138-209 (defaults + vg preset + colinear defaults) and
// :211-262 (the options the hot path honours).  Shared by ref_main.cpp (whole
// program) and ref_dump.cpp (stage-level dumper).
#ifndef GC_ORACLE_REF_PARAMS_H
#define GC_ORACLE_REF_PARAMS_H
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <string>
#include <vector>
#include "Aligner.h"

inline AlignerParams gcDefaultParams()
{
	AlignerParams params;
	params.graphFile = ""; params.outputGAMFile = ""; params.outputJSONFile = ""; params.outputGAFFile = "";
	params.outputCorrectedFile = ""; params.outputCorrectedClippedFile = "";
	params.numThreads = 1; params.initialBandwidth = 0; params.rampBandwidth = 0; params.dynamicRowStart = false;
	params.maxCellsPerSlice = std::numeric_limits<decltype(params.maxCellsPerSlice)>::max();
	params.verboseMode = false; params.shortVerboseMode = false; params.tryAllSeeds = false; params.highMemory = false;
	params.mxmLength = 20; params.mumCount = 0; params.memCount = 0; params.seederCachePrefix = "";
	params.alignmentSelectionMethod = AlignmentSelection::SelectionMethod::GreedyLength;
	params.selectionECutoff = -1; params.forceGlobal = false; params.compressCorrected = false; params.compressClipped = false;
	params.preciseClipping = false; params.minimizerSeedDensity = 0; params.minimizerLength = 19; params.minimizerWindowSize = 30;
	params.seedClusterMinSize = 1; params.minimizerDiscardMostNumerousFraction = 0.0002; params.seedExtendDensity = 0.002;
	params.nondeterministicOptimizations = false; params.optimalDijkstra = false; params.preciseClippingIdentityCutoff = 0.5;
	params.Xdropcutoff = 0; params.DPRestartStride = 0; params.cigarMatchMismatchMerge = false;
	params.colinearChaining = true; params.generatePath = false; params.generatePathSeed = 0; params.IndexMpcFile = "";
	params.fastMode = false; params.graphStatistics = false;
	// vg preset (AlignerMain.cpp:186-193)
	params.minimizerSeedDensity = 10; params.minimizerLength = 15; params.minimizerWindowSize = 20; params.seedExtendDensity = -1;
	params.minimizerDiscardMostNumerousFraction = 0.001; params.nondeterministicOptimizations = false; params.initialBandwidth = 10;
	// colinear defaults (AlignerMain.cpp:201-209)
	params.alignmentSelectionMethod = AlignmentSelection::SelectionMethod::All;
	params.tryAllSeeds = true; params.colinearGap = 10000; params.colinearSplitLen = 35; params.colinearSplitGap = 35; params.samplingStep = 1;
	return params;
}

// returns the list of -a outputs; exits on unknown options like the reference does
inline AlignerParams gcParseArgs(int argc, char** argv, std::vector<std::string>& extra)
{
	AlignerParams params = gcDefaultParams();
	std::vector<std::string> outputAlns;
	bool splitGapGiven = false;
	for (int i = 1; i < argc; i++)
	{
		std::string a = argv[i];
		auto next = [&](void) -> std::string { if (i + 1 >= argc) { std::cerr << "missing value for " << a << std::endl; std::exit(1); } return argv[++i]; };
		if (a == "-g" || a == "--graph") params.graphFile = next();
		else if (a == "-f" || a == "--reads") { params.fastqFiles.push_back(next()); while (i + 1 < argc && argv[i+1][0] != '-') params.fastqFiles.push_back(argv[++i]); }
		else if (a == "-a" || a == "--alignments-out") outputAlns.push_back(next());
		else if (a == "-t" || a == "--threads") params.numThreads = std::stoull(next());
		else if (a == "-b" || a == "--bandwidth") params.initialBandwidth = std::stoull(next());
		else if (a == "--colinear-gap") params.colinearGap = std::stoll(next());
		else if (a == "--colinear-split-len") params.colinearSplitLen = std::stoll(next());
		else if (a == "--colinear-split-gap") { params.colinearSplitGap = std::stoll(next()); splitGapGiven = true; }
		else if (a == "--sampling-step") params.samplingStep = (double)std::stoll(next()); // long long, AlignerMain.cpp:43,233
		else if (a == "--fast-mode") params.fastMode = true;
		else if (a == "--no-colinear-chaining")
		{
			// AlignerMain.cpp:198-209: the colinear defaults are simply not applied
			AlignerParams d = gcDefaultParams();
			params.colinearChaining = false;
			params.alignmentSelectionMethod = AlignmentSelection::SelectionMethod::GreedyLength;
			params.tryAllSeeds = false;
			(void)d;
		}
		else if (a == "--verbose") params.verboseMode = true;
		else if (a == "--short-verbose") params.shortVerboseMode = true;
		else if (a.rfind("--gc-", 0) == 0) { extra.push_back(a); if (i + 1 < argc && argv[i+1][0] != '-') extra.push_back(argv[++i]); }
		else { std::cerr << "unrecognised option '" << a << "'" << std::endl << "run with option -h for help" << std::endl; std::exit(1); }
	}
	if (params.samplingStep != 1)
	{
		if (splitGapGiven) std::cerr << "WARNING: --sampling-step and --colinear-split-gap are both set! --colinear-split-gap will be ignored, and set to (--sampling-step * --colinear-split-len)" << std::endl;
		params.colinearSplitGap = ceil(params.samplingStep * params.colinearSplitLen);
	}
	for (std::string file : outputAlns)
	{
		if (file.size() >= 4 && file.substr(file.size()-4) == ".gam") params.outputGAMFile = file;
		else if (file.size() >= 5 && file.substr(file.size()-5) == ".json") params.outputJSONFile = file;
		else if (file.size() >= 4 && file.substr(file.size()-4) == ".gaf") params.outputGAFFile = file;
		else { std::cerr << "unknown output alignment format (" << file << "), must be either .gaf, .gam or .json" << std::endl; std::exit(1); }
	}
	return params;
}
#endif
