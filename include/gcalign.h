/* libgcalign -- batch form of the reference's per-read loop body for host callers.
 *
 * One gcalign_align() call replaces, for a batch of reads, the body of
 *   runComponentMappings(...)                      (src/Aligner.cpp:461-1065, colinear mode)
 * i.e. everything between pulling a FastQ from the input queue (:499-506) and handing the
 * serialised vg::Alignment records to the writer queue (:1047-1048).  The graph / MPC /
 * minimizer index it needs replaces getGraph + buildMPC + MinimizerSeeder construction
 * (src/Aligner.cpp:1137-1167) and is built by gcalign_open from a .gfa, or loaded from the
 * flat .gcidx form (graphchainer_b200/csrc/gc_index.h).
 * The DP stages run on the GPU through libgcgpu (include/gcgpu.h); there is no CPU path.
 * All buffers are caller-owned host memory.  Returns 0 or a negative gcgpu_status.
 */
#ifndef GCALIGN_H
#define GCALIGN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcalign gcalign;

/* AlignerParams fields the path honours (src/Aligner.h:10-63) */
typedef struct gcalign_options
{
	int32_t device;               /* CUDA device index                                   */
	int32_t host_threads;         /* -t: threads for the host stages                      */
	int32_t initial_bandwidth;    /* -b, default 10                                       */
	int32_t streams;              /* read batches in flight per device, each with its own CUDA stream,
	                                 workspaces and share of the host threads (0 = default 6)     */
	int64_t colinear_gap;         /* --colinear-gap, default 10000                        */
	int64_t colinear_split_len;   /* --colinear-split-len, default 35                     */
	int64_t colinear_split_gap;   /* --colinear-split-gap, default 35                     */
	uint64_t batch_bp;            /* read bases per internal GPU batch (0 = default: one round of equal batches over the streams, 8-26 Mbp each) */
	int32_t gzip_level;           /* zlib level of the GAM gzip members: 0 = default (1, fastest);
	                                 the reference's GzipOutputStream uses 6; decoded records
	                                 are identical at every level                            */
	int32_t threads_per_stream;   /* host threads of each in-flight batch (0 = 2.25 * host_threads / streams); more than
	                                 that share oversubscribes the cores on purpose: a batch waiting for the GPU
	                                 leaves its threads idle                                  */
	int32_t no_colinear_chaining; /* --no-colinear-chaining: the whole-read pass and its GreedyLength selection only
	                                 ("align as in GraphAligner", AlignerMain.cpp:108,198)       */
} gcalign_options;

/* per read: the fields of the reference's --short-verbose line (src/Aligner.cpp:909-915) */
typedef struct gcalign_read_summary
{
	uint32_t num_alignments;
	uint32_t used_chain;          /* 1 = the chained (CLC) alignment was written          */
	uint32_t anchors, chained;
	uint64_t path_bp;
	uint64_t clc_score, long_edit_distance;
	uint64_t gam_offset, gam_size; /* this read's GAM record inside the output buffer     */
} gcalign_read_summary;

typedef struct gcalign_stats
{
	double k1_ms, k2_ms, k3_ms;          /* CUDA-event kernel time per stage               */
	uint64_t k1_items, k1_columns;       /* extensions, 64-row Myers column steps          */
	uint64_t k2_anchors;
	uint64_t k3_items, k3_blocks;        /* NW alignments, 64-row block column steps       */
	uint64_t launches;                   /* kernel launches                                */
	uint64_t s1_rounds;
	uint64_t h2d_bytes, d2h_bytes;       /* bytes libgcgpu copied across PCIe for this call (gcgpu_transfer_bytes) */
	uint64_t seeds_found, seeds_extended;
	double s0_ms;                        /* seeding lookups (gcgpu_seed)                   */
} gcalign_stats;

void gcalign_default_options(gcalign_options* opts);
const char* gcalign_last_error(void);
/* graph_path: *.gfa or *.vg (index built here) or *.gcidx (prebuilt index) */
int gcalign_open(const char* graph_path, const gcalign_options* opts, gcalign** out);
void gcalign_close(gcalign* h);
/* measurement aid: gcgpu_int_peak() of the handle's device (int32 LOP3/IADD3 instructions per second, thread level) */
int gcalign_int_peak(gcalign* h, double* int32_ops_per_s);
/* reads: seqs[seq_offsets[i] .. seq_offsets[i+1]) and names likewise.  If gam_out != NULL the
 * reads' GAM records (one gzip member per read with an alignment) are appended to it.  If they
 * do not fit, GCGPU_ERR_ARG is returned and *gam_used holds the number of bytes the records need
 * (the call can be repeated with a buffer of that size).  The calling thread's OpenMP thread
 * count is restored on return; gcalign_open tunes the process's malloc arenas unless
 * GCALIGN_KEEP_MALLOC is set in the environment.                                            */
int gcalign_align(gcalign* h, const char* seqs, const uint64_t* seq_offsets, const char* names, const uint64_t* name_offsets, uint32_t num_reads,
                  uint8_t* gam_out, uint64_t gam_capacity, uint64_t* gam_used, gcalign_read_summary* summaries, gcalign_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
