// libgcalign: the host pipeline behind a C ABI (include/gcalign.h).
#include <omp.h>
#include <malloc.h>
#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include "../../include/gcalign.h"
#include "gc_pipeline.h"
#include "gc_output.h"
#include "gc_post_host.h"
#include "gc_builder.h"

static thread_local std::string g_alignError;
static int fail(int code, const std::string& msg) { g_alignError = msg; return code; }

// one batch in flight: a libgcgpu context (own stream + workspaces) and a pipeline whose page-locked
// buffers persist across calls
struct GcAlignWorker
{
	gcgpu_ctx* ctx = nullptr;
	std::unique_ptr<GcPipeline> pipeline;
};
struct gcalign
{
	GcHostGraph graph;
	std::vector<GcAlignWorker> workers;
	gcalign_options opts;
	GcPipelineParams pipe;
};

extern "C" void gcalign_default_options(gcalign_options* o)
{
	memset(o, 0, sizeof(*o));
	o->device = 0; o->host_threads = 0; o->initial_bandwidth = 10; o->streams = 0;
	o->colinear_gap = 10000; o->colinear_split_len = 35; o->colinear_split_gap = 35; o->batch_bp = 0;
}
extern "C" const char* gcalign_last_error(void) { return g_alignError.c_str(); }

extern "C" int gcalign_open(const char* graph_path, const gcalign_options* opts, gcalign** out)
{
	if (!graph_path || !out) return fail(GCGPU_ERR_ARG, "gcalign_open: null argument");
	*out = nullptr;
	// The host stages allocate and free gigabytes of per-batch vectors from several threads.  With glibc's defaults every
	// large block is its own mmap/munmap (page faults on first touch, TLB shootdowns across all threads on release) and the
	// arenas are trimmed back to the kernel between batches; keeping the memory in the arenas was worth 7 % end to end on
	// B200 (profiles/r01j).
	// This changes the allocator of the whole process: GCALIGN_KEEP_MALLOC=1 leaves it alone.
	if (!getenv("GCALIGN_KEEP_MALLOC"))
	{
		mallopt(M_MMAP_THRESHOLD, 1 << 30);
		mallopt(M_TRIM_THRESHOLD, -1);
		mallopt(M_TOP_PAD, 256 << 20);
	}
	gcalign* h = new gcalign();
	if (opts) h->opts = *opts; else gcalign_default_options(&h->opts);
	if (h->opts.colinear_split_gap < 1 || h->opts.colinear_split_len < 1) { delete h; return fail(GCGPU_ERR_ARG, "gcalign_open: split length / gap must be >= 1"); }
	try
	{
		std::string path = graph_path;
		GcIndexFile idx;
		if (path.size() > 6 && path.substr(path.size() - 6) == ".gcidx") idx.load(path);
		else idx = gcbuild::buildIndexFromGfa(path, 15, 20, 0.001, false);
		h->graph.fromIndex(idx);
	}
	catch (const std::exception& e) { delete h; return fail(GCGPU_ERR_ARG, std::string("gcalign_open: ") + e.what()); }
	const GcHostGraph& g = h->graph;
	gcgpu_graph gg; memset(&gg, 0, sizeof(gg));
	gg.num_nodes = (uint32_t)g.numNodes();
	gg.node_length = g.nodeLength.data(); gg.node_seq = g.nodeSeq.data();
	gg.in_start = g.inStart.data(); gg.in_nbr = g.inNbr.data(); gg.out_start = g.outStart.data(); gg.out_nbr = g.outNbr.data();
	gg.component_number = g.componentNumber.data(); gg.linearizable = g.linearizable.data();
	gg.num_components = (uint32_t)g.compStart.size() - 1;
	gg.comp_map = g.compMap.data(); gg.comp_idx = g.compIdx.data(); gg.comp_start = g.compStart.data(); gg.topo_ids = g.topoIds.data();
	gg.paths_start = g.pathsStart.data(); gg.paths_k = g.pathsK.data(); gg.back_start = g.backStart.data(); gg.back_node = g.backNode.data(); gg.back_k = g.backK.data();
	gcFillOrigArrays(g, gg);
	gcgpu_params gp; gp.initial_bandwidth = h->opts.initial_bandwidth;
	int streams = h->opts.streams > 0 ? h->opts.streams : 6;
	h->workers.resize(streams);
	for (int w = 0; w < streams; w++)
	{
		int rc = gcgpu_create(h->opts.device, &gg, &gp, &h->workers[w].ctx);
		if (rc == GCGPU_OK) rc = gcUploadMinimizerIndex(h->workers[w].ctx, g);
		if (rc == GCGPU_OK) rc = gcUploadNodeNames(h->workers[w].ctx, g);
		if (rc != GCGPU_OK) { std::string msg = gcgpu_last_error(); gcalign_close(h); return fail(rc, "gcalign_open: " + msg); }
	}
	h->pipe.colinearGap = h->opts.colinear_gap; h->pipe.colinearSplitLen = h->opts.colinear_split_len; h->pipe.colinearSplitGap = h->opts.colinear_split_gap;
	h->pipe.colinearChaining = h->opts.no_colinear_chaining == 0;
	*out = h;
	return GCGPU_OK;
}

extern "C" void gcalign_close(gcalign* h)
{
	if (!h) return;
	for (auto& w : h->workers) { w.pipeline.reset(); if (w.ctx) gcgpu_destroy(w.ctx); }
	delete h;
}

extern "C" int gcalign_int_peak(gcalign* h, double* v)
{
	if (!h || !v || h->workers.empty()) return fail(GCGPU_ERR_ARG, "gcalign_int_peak: null argument");
	int rc = gcgpu_int_peak(h->workers[0].ctx, v);
	if (rc != GCGPU_OK) return fail(rc, std::string("gcalign_int_peak: ") + gcgpu_last_error());
	return GCGPU_OK;
}

extern "C" int gcalign_align(gcalign* h, const char* seqs, const uint64_t* seq_offsets, const char* names, const uint64_t* name_offsets, uint32_t num_reads,
	uint8_t* gam_out, uint64_t gam_capacity, uint64_t* gam_used, gcalign_read_summary* summaries, gcalign_stats* stats)
{
	if (!h || (num_reads && (!seqs || !seq_offsets))) return fail(GCGPU_ERR_ARG, "gcalign_align: null argument");
	if (gam_used) *gam_used = 0;
	if (stats) memset(stats, 0, sizeof(*stats));
	int hostThreads = h->opts.host_threads > 0 ? h->opts.host_threads : omp_get_max_threads();
	// Batches of read bases, handed to the workers in order.  Default: as many equal batches as there are workers (one round: a
	// worker that takes a second batch while the others idle costs more than larger batches gain -- c3, 150 Mbp on 6 workers:
	// 6 x 25 Mbp 167 Mbp/s, 9 x 16.8 Mbp 144, 12 x 12.6 Mbp 137, profiles/r03z), more rounds once a batch would exceed 26 Mbp
	// (device memory: ~0.5 GB of trace slots per Mbp of HiFi reads), fewer batches when they would fall below 8 Mbp (the K1 launches
	// of a batch last as long as their longest read whatever the batch size).  opts.batch_bp sets the batch size outright.
	const uint64_t totalBp = num_reads ? seq_offsets[num_reads] - seq_offsets[0] : 0;
	// Ultra-long reads: every seed extension reserves slabs and a trace slot for the rest of its read, ~4 GB of device buffers per
	// Mbp of 50-100 kb reads (c4; 0.6 GB per Mbp of 10-20 kb reads) -- three batches in flight, not six (profiles/r04b: six ran a
	// 180 GB device out of memory).  Sizing the trace slots after the forward pass would lift this.
	size_t inFlight = std::max<size_t>(1, h->workers.size());
	if (num_reads && totalBp / num_reads > 30000) inFlight = std::min<size_t>(inFlight, 3);
	uint64_t count;
	if (h->opts.batch_bp) count = std::max<uint64_t>(1, (totalBp + h->opts.batch_bp - 1) / h->opts.batch_bp);
	else
	{
		const uint64_t workers = inFlight, largest = 26u << 20, smallest = 8u << 20;
		count = workers * std::max<uint64_t>(1, (totalBp + workers * largest - 1) / (workers * largest));
		if (count == workers && totalBp / workers < smallest) count = std::min<uint64_t>(workers, std::max<uint64_t>(1, (totalBp + smallest / 2) / smallest));
	}
	// batch k ends with the read that reaches k + 1 shares of the bases
	std::vector<std::pair<uint32_t, uint32_t>> batches;
	{
		uint32_t first = 0;
		for (uint64_t k = 0; k < count && first < num_reads; k++)
		{
			const uint64_t upTo = k + 1 == count ? totalBp : (uint64_t)((double)totalBp * (double)(k + 1) / (double)count);
			uint32_t r = first + 1;
			while (r < num_reads && seq_offsets[r] - seq_offsets[0] < upTo) r++;
			if (k + 1 == count) r = num_reads;
			batches.emplace_back(first, r);
			first = r;
		}
	}
	size_t W = std::min(inFlight, std::max<size_t>(1, batches.size()));
	// default: the batches in flight together hold ~2.25x as many threads as there are host threads -- a batch spends more than half
	// of its time waiting for its kernels (threads asleep), measured best on B200 with 16 cores (profiles/r01h: 6 streams x 6 threads)
	int threadsPerWorker = h->opts.threads_per_stream > 0 ? h->opts.threads_per_stream : std::min(hostThreads, std::max(1, (hostThreads * 9 + 4 * (int)W - 1) / (4 * (int)W)));
	std::vector<std::vector<std::string>> records(batches.size());
	std::vector<std::vector<GcReadResult>> allResults(batches.size());
	std::vector<uint64_t> launches0(W), h2d0(W), d2h0(W);
	std::mutex errMutex; std::string error;
	GcPipelineStats total;
	auto tCall0 = std::chrono::steady_clock::now();
	const int callerOmpThreads = omp_get_max_threads(); // worker 0 runs on the caller's thread: its OpenMP setting is restored below
	auto work = [&](size_t w)
	{
		omp_set_num_threads(threadsPerWorker);
		GcAlignWorker& wk = h->workers[w];
		try
		{
			if (!wk.pipeline) wk.pipeline.reset(new GcPipeline(h->graph, wk.ctx, h->pipe));
			GcPipeline& pipeline = *wk.pipeline;
			pipeline.stats = GcPipelineStats();
			std::vector<GcRead> batch;
			// worker w takes batches w, w + W, ...: the batches are equal, and a worker that sees the same share of the same input again
			// (a caller streaming similar calls) finds its context's device buffers already the right size -- with a shared counter any
			// worker got any batch, and a context regrew its multi-GB buffers whenever it met a slightly larger batch than before
			for (size_t bi = w; bi < batches.size(); bi += W)
			{
				{ std::lock_guard<std::mutex> lock(errMutex); if (!error.empty()) break; }
				batch.clear();
				for (uint32_t r = batches[bi].first; r < batches[bi].second; r++)
				{
					GcRead rd;
					rd.sequence.assign(seqs + seq_offsets[r], seq_offsets[r + 1] - seq_offsets[r]);
					if (names && name_offsets) rd.name.assign(names + name_offsets[r], name_offsets[r + 1] - name_offsets[r]);
					else rd.name = "read_" + std::to_string(r);
					batch.push_back(std::move(rd));
				}
				std::vector<GcReadResult>& results = allResults[bi];
				auto tB0 = std::chrono::steady_clock::now();
				// level 1 (the default) = records of the whole-read alignments encoded and compressed on the device; zlib levels stay on the host
				pipeline.setGamOnDevice(gam_out != nullptr && (h->opts.gzip_level <= 1));
				pipeline.alignBatch(batch, results);
				if (getenv("GC_TRACE_CALL")) fprintf(stderr, "[gcalign] worker %zu batch %zu: alignBatch %.1f ms (start +%.1f ms)\n", w, bi, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tB0).count(), std::chrono::duration<double, std::milli>(tB0 - tCall0).count());
				records[bi].resize(batch.size());
				auto tGam0 = std::chrono::steady_clock::now();
				if (gam_out)
				{
					const int level = h->opts.gzip_level > 0 ? h->opts.gzip_level : 1;
					#pragma omp parallel
					{
						gcout::GamEncoder enc;
						#pragma omp for schedule(dynamic, 4)
						for (size_t i = 0; i < batch.size(); i++)
						{
							if (results[i].alignments.empty()) continue;
							if (!results[i].gamRecord.empty()) { records[bi][i].swap(results[i].gamRecord); continue; } // made on the device
							records[bi][i] = gcout::gamRecordDirect(h->graph, batch[i].name, batch[i].sequence, results[i].alignments, level, enc);
						}
					}
				}
				if (getenv("GC_TRACE")) fprintf(stderr, "[gc] phase gam        %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tGam0).count());
				// the edit runs are not needed past this point: keep only what the summaries read
				for (auto& res : results) for (auto& a : res.alignments) { std::vector<uint32_t>().swap(a.tokens); }
			}
			std::lock_guard<std::mutex> lock(errMutex);
			const GcPipelineStats& ps = pipeline.stats;
			total.s0Ms += ps.s0Ms; total.k1Ms += ps.k1Ms; total.k2Ms += ps.k2Ms; total.k3Ms += ps.k3Ms; total.k1Items += ps.k1Items; total.k1Columns += ps.k1Columns; total.k2Anchors += ps.k2Anchors;
			total.k3Items += ps.k3Items; total.k3Blocks += ps.k3Blocks; total.s1Rounds += ps.s1Rounds; total.s1Wasted += ps.s1Wasted;
		}
		catch (const std::exception& e) { std::lock_guard<std::mutex> lock(errMutex); error = e.what(); }
	};
	for (size_t w = 0; w < W; w++) { launches0[w] = gcgpu_launch_count(h->workers[w].ctx); gcgpu_transfer_bytes(h->workers[w].ctx, &h2d0[w], &d2h0[w]); }
	tCall0 = std::chrono::steady_clock::now();
	{
		std::vector<std::thread> threads;
		for (size_t w = 1; w < W; w++) threads.emplace_back(work, w);
		work(0);
		for (auto& t : threads) t.join();
	}
	if (!error.empty()) { omp_set_num_threads(callerOmpThreads); return fail(GCGPU_ERR_INTERNAL, std::string("gcalign_align: ") + error); }
	auto tCall1 = std::chrono::steady_clock::now();
	// where every read's record goes in the caller's buffer (input order), then the copies with all host threads (~15 bytes per read base)
	uint64_t used = 0;
	std::vector<uint64_t> batchOffset(batches.size() + 1, 0);
	for (size_t bi = 0; bi < batches.size(); bi++)
	{
		batchOffset[bi] = used;
		uint32_t first = batches[bi].first;
		for (size_t i = 0; i < allResults[bi].size(); i++)
		{
			const GcReadResult& res = allResults[bi][i];
			if (summaries)
			{
				gcalign_read_summary& s = summaries[first + i];
				s.num_alignments = (uint32_t)res.alignments.size(); s.used_chain = res.usedChain ? 1 : 0; s.anchors = (uint32_t)res.anchors; s.chained = (uint32_t)res.chained;
				s.path_bp = res.pathBp; s.clc_score = res.clcScore; s.long_edit_distance = res.hasLong ? res.longEditDistance : (uint64_t)-1;
				s.gam_offset = used; s.gam_size = records[bi][i].size();
			}
			if (stats) { stats->seeds_found += res.seedsFound; if (!res.alignments.empty()) stats->seeds_extended += res.seedsExtended; }
			used += records[bi][i].size(); // keeps counting past the capacity: on overflow the caller learns the size it needs
		}
	}
	if (gam_out && used <= gam_capacity)
	{
		omp_set_num_threads(hostThreads);
		for (size_t bi = 0; bi < batches.size(); bi++)
		{
			const std::vector<std::string>& recs = records[bi];
			std::vector<uint64_t> at(recs.size());
			uint64_t o = batchOffset[bi];
			for (size_t i = 0; i < recs.size(); i++) { at[i] = o; o += recs[i].size(); }
			#pragma omp parallel for schedule(static)
			for (size_t i = 0; i < recs.size(); i++) if (!recs[i].empty()) memcpy(gam_out + at[i], recs[i].data(), recs[i].size());
		}
	}
	omp_set_num_threads(callerOmpThreads);
	if (stats)
	{
		stats->s0_ms = total.s0Ms; stats->k1_ms = total.k1Ms; stats->k2_ms = total.k2Ms; stats->k3_ms = total.k3Ms;
		stats->k1_items = total.k1Items; stats->k1_columns = total.k1Columns; stats->k2_anchors = total.k2Anchors;
		stats->k3_items = total.k3Items; stats->k3_blocks = total.k3Blocks; stats->s1_rounds = total.s1Rounds;
		for (size_t w = 0; w < W; w++)
		{
			stats->launches += gcgpu_launch_count(h->workers[w].ctx) - launches0[w];
			uint64_t a = 0, b = 0;
			gcgpu_transfer_bytes(h->workers[w].ctx, &a, &b);
			stats->h2d_bytes += a - h2d0[w]; stats->d2h_bytes += b - d2h0[w];
		}
	}
	if (gam_used) *gam_used = used;
	if (gam_out && used > gam_capacity) return fail(GCGPU_ERR_ARG, "gcalign_align: GAM buffer too small, need " + std::to_string(used) + " bytes (*gam_used holds the size)");
	if (getenv("GC_TRACE_CALL"))
	{
		auto tCall2 = std::chrono::steady_clock::now();
		fprintf(stderr, "[gcalign] batches=%zu workers=%zu threads/worker=%d | workers %.1f ms, gather %.1f ms\n", batches.size(), W, threadsPerWorker,
			std::chrono::duration<double, std::milli>(tCall1 - tCall0).count(), std::chrono::duration<double, std::milli>(tCall2 - tCall1).count());
	}
	return GCGPU_OK;
}
