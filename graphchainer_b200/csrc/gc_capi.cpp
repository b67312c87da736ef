// libgcalign: the host pipeline behind a C ABI (include/gcalign.h).
#include <omp.h>
#include <cstring>
#include <memory>
#include <string>
#include "../../include/gcalign.h"
#include "gc_pipeline.h"
#include "gc_output.h"
#include "gc_builder.h"

static thread_local std::string g_alignError;
static int fail(int code, const std::string& msg) { g_alignError = msg; return code; }

struct gcalign
{
	GcHostGraph graph;
	gcgpu_ctx* ctx = nullptr;
	gcalign_options opts;
	GcPipelineParams pipe;
	std::unique_ptr<GcPipeline> pipeline; // persistent: its page-locked buffers are reused by every call
};

extern "C" void gcalign_default_options(gcalign_options* o)
{
	memset(o, 0, sizeof(*o));
	o->device = 0; o->host_threads = 0; o->initial_bandwidth = 10;
	o->colinear_gap = 10000; o->colinear_split_len = 35; o->colinear_split_gap = 35; o->batch_bp = 0;
}
extern "C" const char* gcalign_last_error(void) { return g_alignError.c_str(); }

extern "C" int gcalign_open(const char* graph_path, const gcalign_options* opts, gcalign** out)
{
	if (!graph_path || !out) return fail(GCGPU_ERR_ARG, "gcalign_open: null argument");
	*out = nullptr;
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
	gcgpu_params gp; gp.initial_bandwidth = h->opts.initial_bandwidth;
	int rc = gcgpu_create(h->opts.device, &gg, &gp, &h->ctx);
	if (rc != GCGPU_OK) { std::string msg = gcgpu_last_error(); delete h; return fail(rc, "gcalign_open: " + msg); }
	h->pipe.colinearGap = h->opts.colinear_gap; h->pipe.colinearSplitLen = h->opts.colinear_split_len; h->pipe.colinearSplitGap = h->opts.colinear_split_gap;
	*out = h;
	return GCGPU_OK;
}

extern "C" void gcalign_close(gcalign* h)
{
	if (!h) return;
	h->pipeline.reset();
	if (h->ctx) gcgpu_destroy(h->ctx);
	delete h;
}

extern "C" int gcalign_align(gcalign* h, const char* seqs, const uint64_t* seq_offsets, const char* names, const uint64_t* name_offsets, uint32_t num_reads,
	uint8_t* gam_out, uint64_t gam_capacity, uint64_t* gam_used, gcalign_read_summary* summaries, gcalign_stats* stats)
{
	if (!h || (num_reads && (!seqs || !seq_offsets))) return fail(GCGPU_ERR_ARG, "gcalign_align: null argument");
	if (gam_used) *gam_used = 0;
	if (stats) memset(stats, 0, sizeof(*stats));
	if (h->opts.host_threads > 0) omp_set_num_threads(h->opts.host_threads);
	uint64_t batchBp = h->opts.batch_bp ? h->opts.batch_bp : (8u << 20);
	uint64_t launches0 = gcgpu_launch_count(h->ctx);
	uint64_t used = 0;
	try
	{
		if (!h->pipeline) h->pipeline.reset(new GcPipeline(h->graph, h->ctx, h->pipe));
		GcPipeline& pipeline = *h->pipeline;
		pipeline.stats = GcPipelineStats();
		std::vector<GcRead> batch;
		std::vector<GcReadResult> results;
		for (uint32_t first = 0; first < num_reads; )
		{
			batch.clear();
			uint64_t bp = 0;
			uint32_t r = first;
			for (; r < num_reads && (bp < batchBp || r == first); r++)
			{
				GcRead rd;
				rd.sequence.assign(seqs + seq_offsets[r], seq_offsets[r + 1] - seq_offsets[r]);
				if (names && name_offsets) rd.name.assign(names + name_offsets[r], name_offsets[r + 1] - name_offsets[r]);
				else rd.name = "read_" + std::to_string(r);
				bp += rd.sequence.size();
				batch.push_back(std::move(rd));
			}
			pipeline.alignBatch(batch, results);
			std::vector<std::string> records(batch.size());
			if (gam_out)
			{
				#pragma omp parallel for schedule(dynamic, 4)
				for (size_t i = 0; i < batch.size(); i++)
				{
					if (results[i].alignments.empty()) continue;
					std::vector<gcout::Alignment> alns;
					for (const GcAlnItem& item : results[i].alignments) alns.push_back(gcout::toAlignment(h->graph, batch[i].name, batch[i].sequence, item));
					records[i] = gcout::gamRecord(alns);
				}
			}
			for (size_t i = 0; i < batch.size(); i++)
			{
				const GcReadResult& res = results[i];
				if (summaries)
				{
					gcalign_read_summary& s = summaries[first + i];
					s.num_alignments = (uint32_t)res.alignments.size(); s.used_chain = res.usedChain ? 1 : 0; s.anchors = (uint32_t)res.anchors; s.chained = (uint32_t)res.chained;
					s.path_bp = res.pathBp; s.clc_score = res.clcScore; s.long_edit_distance = res.hasLong ? res.longEditDistance : (uint64_t)-1;
					s.gam_offset = used; s.gam_size = records[i].size();
				}
				if (stats) { stats->seeds_found += res.seedsFound; if (!res.alignments.empty()) stats->seeds_extended += res.seedsExtended; }
				if (gam_out && !records[i].empty())
				{
					if (used + records[i].size() > gam_capacity) return fail(GCGPU_ERR_ARG, "gcalign_align: GAM buffer too small");
					memcpy(gam_out + used, records[i].data(), records[i].size());
					used += records[i].size();
				}
			}
			first = r;
		}
		if (stats)
		{
			stats->k1_ms = pipeline.stats.k1Ms; stats->k2_ms = pipeline.stats.k2Ms; stats->k3_ms = pipeline.stats.k3Ms;
			stats->k1_items = pipeline.stats.k1Items; stats->k1_columns = pipeline.stats.k1Columns; stats->k2_anchors = pipeline.stats.k2Anchors;
			stats->k3_items = pipeline.stats.k3Items; stats->k3_blocks = pipeline.stats.k3Blocks; stats->s1_rounds = pipeline.stats.s1Rounds;
			stats->launches = gcgpu_launch_count(h->ctx) - launches0;
		}
	}
	catch (const std::exception& e) { return fail(GCGPU_ERR_INTERNAL, std::string("gcalign_align: ") + e.what()); }
	if (gam_used) *gam_used = used;
	return GCGPU_OK;
}
