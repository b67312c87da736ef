// GraphChainerB200 -- host driver with the reference's command line and outputs
// (src/AlignerMain.cpp options -g/-f/-a/-t, --colinear-split-len/--colinear-split-gap/
// --sampling-step/--colinear-gap, --short-verbose; GAM/JSON output, src/Aligner.cpp:1124-1310),
// running the per-read hot path on B200 GPUs through libgcgpu.
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <mutex>
#include <queue>
#include <sstream>
#include <thread>
#include <omp.h>
#include <malloc.h>
#include <zlib.h>
#include "gc_pipeline.h"
#include "gc_output.h"
#include "gc_post_host.h"
#ifdef GC_HAVE_BUILDER
#include "gc_builder.h"
#endif

struct DriverParams
{
	std::string graphFile, indexFile, saveIndexFile;
	std::vector<std::string> readFiles;
	std::string outGam, outJson, outGaf;
	bool cigarMatchMismatchMerge = false;
	size_t threads = 1;
	int gpus = 1;
	int streams = 6;
	int gzipLevel = 1;
	int firstDevice = 0;
	int bandwidth = 10;
	bool shortVerbose = false;
	bool quiet = false;
	size_t batchBp = 16u << 20;
	size_t maxReads = (size_t)-1;
	GcPipelineParams pipe;
	double samplingStep = 1;
};

static void usage()
{
	std::cerr <<
		"Mandatory parameters:\n"
		"  -g [ --graph ] arg            input graph (.gfa / .vg), or a prebuilt index with --gc-index\n"
		"  -f [ --reads ] arg            input reads (fasta or fastq, uncompressed or gzipped)\n"
		"  -a [ --alignments-out ] arg   output alignment file (.gaf/.gam/.json)\n"
		"  --cigar-match-mismatch        use M for matches and mismatches in the GAF cigar instead of = and X\n"
		"Colinear chaining parameters:\n"
		"  --sampling-step arg           Sampling step factor (default 1)\n"
		"  --colinear-split-len arg      fragment length [default 35]\n"
		"  --colinear-split-gap arg      distance between consecutive fragments [default 35]\n"
		"  --colinear-gap arg            split the path if consecutive anchors are further apart [default 10000]\n"
		"General parameters:\n"
		"  -t [ --threads ] arg          host threads\n"
		"  -b [ --bandwidth ] arg        alignment bandwidth [default 10]\n"
		"  --short-verbose               print the per-read progress line\n"
		"  --no-colinear-chaining        do not run colinear chaining and align as in GraphAligner, default parameters\n"
		"B200 parameters:\n"
		"  --gc-gpus N                   GPUs to use (reads are partitioned in length-balanced batches)\n"
		"  --gc-streams N                read batches in flight per GPU [default 6]\n"
		"  --gc-gzip-level N             zlib level of the GAM gzip members [default 1; the reference's library default is 6]\n"
		"  --gc-index file.gcidx         load a prebuilt graph/MPC/minimizer index\n"
		"  --gc-save-index file.gcidx    store the index built from -g\n"
		"  --gc-batch-bp N               read bases per GPU batch [default 8388608]\n";
}

static DriverParams parseArgs(int argc, char** argv)
{
	DriverParams p;
	std::vector<std::string> outputAlns;
	bool splitGapGiven = false;
	for (int i = 1; i < argc; i++)
	{
		std::string a = argv[i];
		auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "the required argument for option '" << a << "' is missing" << std::endl << "run with option -h for help" << std::endl; std::exit(1); } return argv[++i]; };
		if (a == "-h" || a == "--help") { usage(); std::exit(0); }
		else if (a == "-g" || a == "--graph") p.graphFile = next();
		else if (a == "-f" || a == "--reads") { p.readFiles.push_back(next()); while (i + 1 < argc && argv[i + 1][0] != '-') p.readFiles.push_back(argv[++i]); }
		else if (a == "-a" || a == "--alignments-out") outputAlns.push_back(next());
		else if (a == "-t" || a == "--threads") p.threads = std::stoull(next());
		else if (a == "-b" || a == "--bandwidth") p.bandwidth = std::stoi(next());
		else if (a == "--colinear-gap") p.pipe.colinearGap = std::stoll(next());
		else if (a == "--colinear-split-len") p.pipe.colinearSplitLen = std::stoll(next());
		else if (a == "--colinear-split-gap") { p.pipe.colinearSplitGap = std::stoll(next()); splitGapGiven = true; }
		else if (a == "--sampling-step") p.samplingStep = std::stod(next()); // README contract: a double (the reference parses long long)
		else if (a == "--short-verbose") { p.shortVerbose = true; p.pipe.exactProgressLine = true; }
		else if (a == "--no-colinear-chaining") p.pipe.colinearChaining = false;
		else if (a == "--cigar-match-mismatch") p.cigarMatchMismatchMerge = true;
		else if (a == "--gc-gpus") p.gpus = std::stoi(next());
		else if (a == "--gc-gzip-level") p.gzipLevel = std::min(9, std::max(1, std::stoi(next())));
		else if (a == "--gc-streams") p.streams = std::max(1, std::stoi(next()));
		else if (a == "--gc-device") p.firstDevice = std::stoi(next());
		else if (a == "--gc-index") p.indexFile = next();
		else if (a == "--gc-save-index") p.saveIndexFile = next();
		else if (a == "--gc-batch-bp") p.batchBp = std::stoull(next());
		else if (a == "--gc-max-reads") p.maxReads = std::stoull(next());
		else if (a == "--gc-quiet") p.quiet = true;
		else { std::cerr << "unrecognised option '" << a << "'" << std::endl << "run with option -h for help" << std::endl; std::exit(1); }
	}
	bool paramError = false;
	if (p.samplingStep != 1)
	{
		if (splitGapGiven) std::cerr << "WARNING: --sampling-step and --colinear-split-gap are both set! --colinear-split-gap will be ignored, and set to (--sampling-step * --colinear-split-len)" << std::endl;
		p.pipe.colinearSplitGap = (long long)ceil(p.samplingStep * p.pipe.colinearSplitLen);
	}
	if (p.pipe.colinearSplitGap < 1) { std::cerr << "--colinear-split-gap must be >= 1" << std::endl; paramError = true; }
	if (p.graphFile == "" && p.indexFile == "") { std::cerr << "graph file must be given" << std::endl; paramError = true; }
	if (p.readFiles.size() == 0) { std::cerr << "read file must be given" << std::endl; paramError = true; }
	if (outputAlns.size() == 0) { std::cerr << "alignments-out must be given" << std::endl; paramError = true; }
	for (std::string file : outputAlns)
	{
		if (file.size() >= 4 && file.substr(file.size() - 4) == ".gam") p.outGam = file;
		else if (file.size() >= 5 && file.substr(file.size() - 5) == ".json") p.outJson = file;
		else if (file.size() >= 4 && file.substr(file.size() - 4) == ".gaf") p.outGaf = file;
		else { std::cerr << "unknown output alignment format (" << file << "), must be either .gaf, .gam or .json" << std::endl; paramError = true; }
	}
	if (p.threads < 1) { std::cerr << "number of threads must be >= 1" << std::endl; paramError = true; }
	if (p.bandwidth < 1) { std::cerr << "default bandwidth must be >= 1" << std::endl; paramError = true; }
	if (paramError) { std::cerr << "run with option -h for help" << std::endl; std::exit(1); }
	return p;
}

// fastqloader.h:10-139 (zlib's gzread handles plain files too)
struct ReadStream
{
	gzFile f = nullptr;
	std::string pending;
	bool havePending = false;
	bool fastq = false;
	bool open(const std::string& filename)
	{
		std::string name = filename;
		if (name.size() > 3 && name.substr(name.size() - 3) == ".gz") name = name.substr(0, name.size() - 3);
		fastq = (name.size() > 6 && name.substr(name.size() - 6) == ".fastq") || (name.size() > 3 && name.substr(name.size() - 3) == ".fq");
		f = gzopen(filename.c_str(), "rb");
		if (f) gzbuffer(f, 1 << 20);
		return f != nullptr;
	}
	bool getline(std::string& line)
	{
		line.clear();
		char buf[1 << 16];
		bool got = false;
		while (gzgets(f, buf, sizeof(buf)))
		{
			got = true;
			size_t n = strlen(buf);
			if (n && buf[n - 1] == '\n') { line.append(buf, n - 1); return true; }
			line.append(buf, n);
		}
		return got;
	}
	bool next(GcRead& read)
	{
		std::string line;
		if (fastq)
		{
			while (getline(line))
			{
				if (line.size() == 0 || line[0] != '@') continue;
				if (line.back() == '\r') line.pop_back();
				read.name = line.substr(1);
				getline(line); if (!line.empty() && line.back() == '\r') line.pop_back();
				read.sequence = line;
				getline(line); getline(line);
				return true;
			}
			return false;
		}
		if (!havePending) { while (getline(line)) if (line.size() && line[0] == '>') { pending = line; havePending = true; break; } }
		if (!havePending) return false;
		if (pending.back() == '\r') pending.pop_back();
		read.name = pending.substr(1);
		read.sequence.clear();
		havePending = false;
		while (getline(line))
		{
			if (line.size() == 0) continue;
			if (line[0] == '>') { pending = line; havePending = true; break; }
			if (line.back() == '\r') line.pop_back();
			read.sequence += line;
		}
		return true;
	}
	void close() { if (f) gzclose(f); f = nullptr; }
};

int main(int argc, char** argv)
{
	DriverParams params = parseArgs(argc, argv);
	mallopt(M_MMAP_THRESHOLD, 1 << 30); mallopt(M_TRIM_THRESHOLD, -1); mallopt(M_TOP_PAD, 256 << 20); // keep the per-batch vectors in the arenas (see gc_capi.cpp)
	omp_set_num_threads((int)params.threads);
	std::cout << "Co-linear chaining " << (params.pipe.colinearChaining ? "on" : "off"); // Aligner.cpp:1127-1132
	if (params.pipe.colinearChaining) std::cout << " splits=(" << params.pipe.colinearSplitLen << "," << params.pipe.colinearSplitGap << "," << params.pipe.colinearGap << ")";
	std::cout << std::endl;
	GcHostGraph graph;
	auto t0 = std::chrono::steady_clock::now();
	try
	{
		if (params.indexFile != "")
		{
			std::cout << "Load index from " << params.indexFile << std::endl;
			GcIndexFile idx; idx.load(params.indexFile);
			graph.fromIndex(idx);
		}
		else
		{
#ifdef GC_HAVE_BUILDER
			std::cout << "Load graph from " << params.graphFile << std::endl;
			GcIndexFile idx = gcbuild::buildIndexFromGfa(params.graphFile, 15, 20, 0.001, !params.quiet);
			if (params.saveIndexFile != "") idx.save(params.saveIndexFile);
			graph.fromIndex(idx);
#else
			std::cerr << "this build has no graph builder: pass --gc-index" << std::endl;
			return 1;
#endif
		}
	}
	catch (const std::exception& e) { std::cerr << "Error in the graph: " << e.what() << std::endl; return 1; }
	double indexSec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::cout << "Index ready in " << indexSec << " s (" << graph.numNodes() << " split nodes)" << std::endl;
	std::cout << "Minimizer seeds, length " << graph.mzLength << ", window size " << graph.mzWindow << ", density " << params.pipe.minimizerSeedDensity << std::endl;
	std::cout << "Seed cluster size " << params.pipe.seedClusterMinSize << std::endl;
	std::cout << "Initial bandwidth " << params.bandwidth << std::endl;
	if (params.outGam != "") std::cout << "write alignments to " << params.outGam << std::endl;
	if (params.outJson != "") std::cout << "write alignments to " << params.outJson << std::endl;
	if (params.outGaf != "") std::cout << "write alignments to " << params.outGaf << std::endl;

	// ---- one libgcgpu context (graph replica) per GPU
	gcgpu_graph gg; memset(&gg, 0, sizeof(gg));
	gg.num_nodes = (uint32_t)graph.numNodes();
	gg.node_length = graph.nodeLength.data(); gg.node_seq = graph.nodeSeq.data();
	gg.in_start = graph.inStart.data(); gg.in_nbr = graph.inNbr.data(); gg.out_start = graph.outStart.data(); gg.out_nbr = graph.outNbr.data();
	gg.component_number = graph.componentNumber.data(); gg.linearizable = graph.linearizable.data();
	gg.num_components = (uint32_t)graph.compStart.size() - 1;
	gg.comp_map = graph.compMap.data(); gg.comp_idx = graph.compIdx.data(); gg.comp_start = graph.compStart.data(); gg.topo_ids = graph.topoIds.data();
	gg.paths_start = graph.pathsStart.data(); gg.paths_k = graph.pathsK.data(); gg.back_start = graph.backStart.data(); gg.back_node = graph.backNode.data(); gg.back_k = graph.backK.data();
	gcFillOrigArrays(graph, gg);
	gcgpu_params gp; gp.initial_bandwidth = params.bandwidth;
	std::vector<gcgpu_ctx*> ctxs;
	// one context (stream + workspaces) per batch in flight: worker w runs on device w % gpus
	const int numWorkers = params.gpus * params.streams;
	for (int d = 0; d < numWorkers; d++)
	{
		gcgpu_ctx* ctx = nullptr;
		if (gcgpu_create(params.firstDevice + d % params.gpus, &gg, &gp, &ctx) != GCGPU_OK) { std::cerr << "gcgpu_create(device " << params.firstDevice + d << ") failed: " << gcgpu_last_error() << std::endl; return 1; }
		if (gcUploadMinimizerIndex(ctx, graph) != GCGPU_OK) { std::cerr << "gcgpu_set_minimizer_index failed: " << gcgpu_last_error() << std::endl; return 1; }
		if (gcUploadNodeNames(ctx, graph) != GCGPU_OK) { std::cerr << "gcgpu_set_node_names failed: " << gcgpu_last_error() << std::endl; return 1; }
		ctxs.push_back(ctx);
	}

	std::ofstream gamOut, jsonOut;
	if (params.outGam != "") gamOut.open(params.outGam, std::ios::binary);
	if (params.outJson != "") jsonOut.open(params.outJson);
	std::ofstream gafOut;
	if (params.outGaf != "") gafOut.open(params.outGaf);
	if ((params.outGam != "" && !gamOut.good()) || (params.outJson != "" && !jsonOut.good()) || (params.outGaf != "" && !gafOut.good())) { std::cerr << "cannot open the alignment output file for writing" << std::endl; return 1; }
	std::mutex outMutex, inMutex;
	bool wroteAny = false;
	size_t statAllAlns = 0; // stats.allAlignmentsCount: alignments before the selection of --no-colinear-chaining, the written ones otherwise (Aligner.cpp:924)
	size_t statReads = 0, statBp = 0, statSeedsFound = 0, statSeedsExtended = 0, statReadsWithSeed = 0, statBpWithSeed = 0, statReadsWithAln = 0, statAlns = 0, statBpAln = 0, statFull = 0, statBpFull = 0;
	size_t readCounter = 0;
	bool anyDropped = false;
	GcPipelineStats total;

	std::cout << "Align" << std::endl;
	auto alignStart = std::chrono::steady_clock::now();
	// reader state shared by the device workers
	size_t fileIdx = 0;
	ReadStream rs; bool rsOpen = false;
	size_t readsTaken = 0;
	auto nextBatch = [&](std::vector<GcRead>& batch)
	{
		std::lock_guard<std::mutex> lock(inMutex);
		batch.clear();
		size_t bp = 0;
		while (bp < params.batchBp && readsTaken < params.maxReads)
		{
			if (!rsOpen)
			{
				if (fileIdx >= params.readFiles.size()) break;
				if (!rs.open(params.readFiles[fileIdx])) { std::cerr << "cannot open " << params.readFiles[fileIdx] << std::endl; fileIdx++; continue; }
				rsOpen = true;
			}
			GcRead r;
			if (!rs.next(r)) { rs.close(); rsOpen = false; fileIdx++; continue; }
			bp += r.sequence.size();
			batch.push_back(std::move(r));
			readsTaken++;
		}
		return !batch.empty();
	};
	auto worker = [&](int d)
	{
		// the batches in flight together hold ~2.25x the host threads: a batch waiting for its kernels leaves its threads asleep
		omp_set_num_threads(std::min<int>((int)params.threads, std::max<int>(1, ((int)params.threads * 9 + 4 * numWorkers - 1) / (4 * numWorkers))));
		GcPipeline pipeline(graph, ctxs[d], params.pipe);
		// GAM only, at the driver's own compression level: the records of the whole-read alignments are made on the device
		pipeline.setGamOnDevice(params.outGam != "" && params.outJson == "" && params.outGaf == "" && params.gzipLevel == 1);
		std::vector<GcRead> batch;
		std::vector<GcReadResult> results;
		while (nextBatch(batch))
		{
			// length-balanced: longest reads first inside a batch (the kernels sort their work items the same way)
			pipeline.alignBatch(batch, results);
			std::vector<std::string> gamRecords(batch.size()), jsonRecords(batch.size()), gafRecords(batch.size());
			auto tGam0 = std::chrono::steady_clock::now();
			#pragma omp parallel
			{
				gcout::GamEncoder enc;
				#pragma omp for schedule(dynamic, 4)
				for (size_t r = 0; r < batch.size(); r++)
				{
					if (results[r].alignments.empty()) continue;
					if (params.outGam != "") { if (!results[r].gamRecord.empty()) gamRecords[r].swap(results[r].gamRecord); else gamRecords[r] = gcout::gamRecordDirect(graph, batch[r].name, batch[r].sequence, results[r].alignments, params.gzipLevel, enc); }
					if (params.outJson != "")
						for (const GcAlnItem& item : results[r].alignments) { jsonRecords[r] += gcout::jsonLine(gcout::toAlignment(graph, batch[r].name, batch[r].sequence, item)); jsonRecords[r] += '\n'; }
					if (params.outGaf != "")
						for (const GcAlnItem& item : results[r].alignments) { gafRecords[r] += gcout::gafLine(graph, batch[r].name, batch[r].sequence, item, params.cigarMatchMismatchMerge); gafRecords[r] += '\n'; }
				}
			}
			if (getenv("GC_TRACE")) fprintf(stderr, "[gc] phase gam        %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tGam0).count());
			std::lock_guard<std::mutex> lock(outMutex);
			for (size_t r = 0; r < batch.size(); r++)
			{
				const GcReadResult& res = results[r];
				statReads++; statBp += batch[r].sequence.size();
				readCounter++;
				statSeedsFound += res.seedsFound;
				statAllAlns += params.pipe.colinearChaining ? res.alignments.size() : res.alignmentsBeforeSelection;
				if (res.seedsFound) { const size_t calls = params.pipe.colinearChaining ? 2 : 1; statReadsWithSeed += calls; statBpWithSeed += calls * batch[r].sequence.size(); } // once per getSeeds call
				if (res.dropped || res.broke) anyDropped = true;
				if (params.shortVerbose && params.pipe.colinearChaining && res.seedsFound && !res.dropped) // the line is printed by the chaining branch only (Aligner.cpp:909-915)
				{
					std::string short_id;
					for (char c : batch[r].name) { if (isspace(c)) break; short_id += c; }
					// the reference's line (Aligner.cpp:909-915); its three stage times are per-read wall clocks that a batched pipeline does
					// not have (printed as 0), and it prints an uninitialised long_edit_distance when the read has no whole-read alignment
					std::cerr << readCounter << " " << short_id << " len=" << batch[r].sequence.length() << " : "
						<< "chained " << res.chained << " / " << res.anchors << " anchors, actual " << res.pathBp << " bps, "
						<< "time 0 0 0  "
						<< "score=" << res.clcScore
						<< " long_edit_distance=" << (res.hasLong ? res.longEditDistance : batch[r].sequence.length())
						<< " one_node_overlaps=" << res.oneNodeOverlapsNow << " / " << res.oneNodeOverlapsAll << std::endl;
				}
				if (res.dropped) std::cerr << "Read " << batch[r].name << " alignment failed (assertion!)" << std::endl;
				if (res.alignments.empty()) continue;
				statSeedsExtended += res.seedsExtended;
				statReadsWithAln++;
				for (const GcAlnItem& a : res.alignments)
				{
					statAlns++;
					size_t sz = a.alignmentEnd - a.alignmentStart;
					if (sz == batch[r].sequence.size()) { statFull++; statBpFull += sz; }
					statBpAln += sz;
				}
				if (params.outGam != "") { gamOut.write(gamRecords[r].data(), gamRecords[r].size()); wroteAny = true; }
				if (params.outJson != "") jsonOut << jsonRecords[r];
				if (params.outGaf != "") gafOut << gafRecords[r];
			}
			total.k1Items += pipeline.stats.k1Items; total.k1Columns += pipeline.stats.k1Columns; total.k1Ms += pipeline.stats.k1Ms; total.k1Launches += pipeline.stats.k1Launches;
			total.s0Ms += pipeline.stats.s0Ms; total.k2Ms += pipeline.stats.k2Ms; total.k2Anchors += pipeline.stats.k2Anchors; total.k3Ms += pipeline.stats.k3Ms; total.k3Items += pipeline.stats.k3Items; total.k3Blocks += pipeline.stats.k3Blocks;
			total.s1Rounds += pipeline.stats.s1Rounds;
			pipeline.stats = GcPipelineStats();
		}
	};
	std::vector<std::thread> workers;
	std::string workerError;
	for (int d = 0; d < numWorkers; d++) workers.emplace_back([&, d]() { try { worker(d); } catch (const std::exception& e) { std::lock_guard<std::mutex> lock(outMutex); workerError = e.what(); } });
	for (auto& w : workers) w.join();
	if (!workerError.empty()) { std::cerr << "fatal: " << workerError << std::endl; return 1; }
	if (params.outGam != "" && !wroteAny) { std::string empty = gcout::gamRecord({}); gamOut.write(empty.data(), empty.size()); } // Aligner.cpp:228-240
#ifdef GC_PROF
	if (getenv("GC_TRACE")) for (int i = 13; i < 32; i++) if (g_profName[i]) fprintf(stderr, "[prof] %-32s %.2f ms\n", g_profName[i], g_prof[i] / 2.0e6);
#endif
	double alignSec = std::chrono::duration<double>(std::chrono::steady_clock::now() - alignStart).count();
	for (auto ctx : ctxs) gcgpu_destroy(ctx);

	std::cout << "Alignment finished" << std::endl;
	std::cout << "Input reads: " << statReads << " (" << statBp << "bp)" << std::endl;
	std::cout << "Seeds found: " << statSeedsFound << std::endl;
	std::cout << "Seeds extended: " << statSeedsExtended << std::endl;
	std::cout << "Reads with a seed: " << statReadsWithSeed << " (" << statBpWithSeed << "bp)" << std::endl;
	std::cout << "Reads with an alignment: " << statReadsWithAln << std::endl;
	std::cout << "Alignments: " << statAlns << " (" << statBpAln << "bp)";
	if (statAllAlns > statAlns) std::cout << " (" << (statAllAlns - statAlns) << " additional alignments discarded)"; // Aligner.cpp:1303
	std::cout << std::endl;
	std::cout << "End-to-end alignments: " << statFull << " (" << statBpFull << "bp)" << std::endl;
	if (anyDropped) std::cout << "Alignment broke with some reads. Look at stderr output." << std::endl;
	gamOut.flush(); jsonOut.flush(); gafOut.flush();
	if ((params.outGam != "" && !gamOut.good()) || (params.outJson != "" && !jsonOut.good()) || (params.outGaf != "" && !gafOut.good())) { std::cerr << "writing the alignment output failed" << std::endl; return 1; }
	if (!params.quiet)
	{
		std::cout << "B200: align phase " << alignSec << " s, " << (statBp / alignSec) << " bp/s on " << params.gpus << " GPU(s); K1 " << total.k1Items << " extensions / " << total.k1Columns << " column steps / " << total.k1Ms << " ms, "
			<< "S0 " << total.s0Ms << " ms, K2 " << total.k2Anchors << " anchors / " << total.k2Ms << " ms, K3 " << total.k3Items << " alignments / " << total.k3Blocks << " block steps / " << total.k3Ms << " ms, S1 rounds " << total.s1Rounds << std::endl;
	}
	return 0;
}
