/* libgcgpu -- C ABI of the B200 (sm_100a) GraphChainer alignment hot path.
 *
 * Drop-in boundary: the reference reaches this path by direct
 * C++ calls from the per-read loop body runComponentMappings (src/Aligner.cpp:492-1062).
 * Each entry point below is the batch form of one of those seams; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers (pinned if the
 * caller wants asynchronous copies), no exceptions across the boundary -- every call
 * returns 0 on success or a negative gcgpu_status, with text in gcgpu_last_error().
 * One gcgpu_ctx per device; calls on one ctx are serialised by the caller.
 * There is no CPU fallback: without a CUDA device gcgpu_create() fails.
 */
#ifndef GCGPU_H
#define GCGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcgpu_ctx gcgpu_ctx;

enum gcgpu_status
{
	GCGPU_OK = 0,
	GCGPU_ERR_CUDA = -1,      /* a CUDA runtime call failed                       */
	GCGPU_ERR_ARG = -2,       /* invalid argument                                 */
	GCGPU_ERR_NOMEM = -3,     /* device or host memory exhausted                  */
	GCGPU_ERR_INTERNAL = -4   /* a work item hit a state the reference asserts on */
};

/* per work item outcome (K1), mirrors OnewayTrace::failed() (GraphAlignerCommon.h:164-171) */
enum gcgpu_item_status
{
	GCGPU_ITEM_OK = 0,
	GCGPU_ITEM_FAILED = 1,    /* the reference's TraceFailed(): <= 1 slice survived        */
	GCGPU_ITEM_INTERNAL = 5   /* assert-class condition; the reference would drop the read */
};

/* Split-node alignment graph + MPC index: flat arrays in REFERENCE numbering, i.e. the
 * members of AlignmentGraph (src/AlignmentGraph.h:145-172) after Finalize() and buildMPC():
 *   node_length[i]      = nodeLength[i]              (1..64, SPLIT_NODE_SIZE, AlignmentGraph.h:20)
 *   node_seq[2i..2i+1]  = nodeSequences[i].s[0..1]   (2 bit/base A0 C1 G2 T3, AlignmentGraph.cpp:114-142)
 *   in_/out_ CSR        = inNeighbors / outNeighbors in stored (insertion) order
 *   component_number    = componentNumber            (AlignmentGraph.cpp:1008-1105)
 *   linearizable        = linearizable               (AlignmentGraph.cpp:644-736)
 *   comp_map/comp_idx   = component_map / component_idx           (AlignmentGraph.cpp:1430-1463)
 *   comp_start          = prefix sums of component_ids[c].size()  (C+1 entries)
 *   topo_ids            = topo_ids[c][idx] at comp_start[c]+idx   (AlignmentGraph.cpp:1351-1364)
 *   paths_* / back_*    = CSR over (comp_start[c]+idx) of paths[c][idx] and backwards[c][idx]
 * All host pointers; gcgpu_create() copies them to the device.  Graphs with ambiguous
 * bases (firstAmbiguous < N) are not supported by this path.                              */
typedef struct gcgpu_graph
{
	uint32_t num_nodes;
	const uint8_t* node_length;
	const uint64_t* node_seq;
	const uint32_t* in_start;
	const uint32_t* in_nbr;
	const uint32_t* out_start;
	const uint32_t* out_nbr;
	const uint32_t* component_number;
	const uint8_t* linearizable;
	/* MPC index (may be all NULL / 0 if gcgpu_chain is not used) */
	uint32_t num_components;
	const uint32_t* comp_map;
	const uint32_t* comp_idx;
	const uint32_t* comp_start;
	const uint32_t* topo_ids;
	const uint32_t* paths_start;
	const uint32_t* paths_k;
	const uint32_t* back_start;
	const uint32_t* back_node;
	const uint32_t* back_k;
	/* original (bigraph-doubled) nodes, needed by the resident batch entry points below (may be NULL / 0 otherwise):
	 *   node_ids[i]    = nodeIDs[i]     digraph node id 2*id+strand of split node i   (AlignmentGraph.h:148)
	 *   node_offset[i] = nodeOffset[i]  offset of split node i in its original node   (AlignmentGraph.h:147)
	 *   orig_*         = nodeLookup + originalNodeSize (AlignmentGraph.h:146,149) as CSR: original node o has digraph id
	 *                    orig_ids[o], length orig_size[o] and the split nodes orig_nodes[orig_start[o] .. orig_start[o+1])
	 *                    in offset order                                                                              */
	const int32_t* node_ids;
	const uint32_t* node_offset;
	uint32_t num_orig;
	const int32_t* orig_ids;
	const uint32_t* orig_start;
	const uint32_t* orig_nodes;
	const uint32_t* orig_size;
} gcgpu_graph;

/* The constants of GraphAlignerCommon::Params the path honours (GraphAlignerCommon.h:92-126);
 * defaults = AlignerMain.cpp:145-209 (vg preset + colinear defaults).                     */
typedef struct gcgpu_params
{
	int32_t initial_bandwidth;   /* -b / params.initialBandwidth, default 10 */
} gcgpu_params;

int gcgpu_version(void);
const char* gcgpu_last_error(void);

/* Replaces: AlignmentGraph construction hand-over + AlignerGraphsizedState (Aligner.cpp:469). */
int gcgpu_create(int device, const gcgpu_graph* graph, const gcgpu_params* params, gcgpu_ctx** out);
void gcgpu_destroy(gcgpu_ctx* ctx);

/* ---- K1: banded bit-parallel graph extension ------------------------------------------
 * One item replaces one call of
 *   GraphAlignerBitvectorBanded::getReverseTraceFromSeed(sequence, bigraphNodeId, nodeOffset, ...)
 * (src/GraphAlignerBitvectorBanded.h:46-71), reached from AlignOneWay -> getAlignmentFromSeed
 * -> getTwoDirectionalTrace (src/GraphAligner.h:114-203,480-525,567-626).
 * `seq` = concatenated IUPAC-mask codes (bit0 A, bit1 C, bit2 G, bit3 T), one byte per base;
 * an item aligns seq[seq_offset .. seq_offset+seq_len) starting from the cell (node, offset)
 * where node is a SPLIT node index (AlignmentGraph::GetUnitigNode already applied).       */
typedef struct gcgpu_ext_item
{
	uint64_t seq_offset;
	int32_t seq_len;
	uint32_t node;
	uint32_t offset;
	uint32_t reserved;
} gcgpu_ext_item;

typedef struct gcgpu_ext_result
{
	int32_t status;         /* gcgpu_item_status */
	int32_t score;          /* OnewayTrace::score */
	uint32_t trace_len;     /* number of trace entries */
	uint32_t reserved;
	uint64_t trace_offset;  /* index of the first entry in the trace buffer */
	uint64_t columns;       /* work counter: 64-row Myers column steps (BVCommon.h:1162 semantics) */
} gcgpu_ext_result;

/* Trace entries are in the order the reference builds them (end of alignment first, the
 * seed cell with seqPos -1 last; getReverseTraceFromTable, BVCommon.h:392-544), packed:
 *   bits 0-31 split node | 32-37 offset in node | 38 nodeSwitch | 39-63 seqPos+1          */
#define GCGPU_TRACE_NODE(t) ((uint32_t)((t) & 0xFFFFFFFFu))
#define GCGPU_TRACE_OFFSET(t) ((uint32_t)(((t) >> 32) & 63))
#define GCGPU_TRACE_SWITCH(t) ((int)(((t) >> 38) & 1))
#define GCGPU_TRACE_SEQPOS(t) ((int32_t)(((t) >> 39) & 0x1FFFFFF) - 1)

/* `seq` may be NULL: the sequence buffer uploaded by the previous gcgpu_extend call on this ctx is
 * reused (a batch of reads is extended in several calls; its codes cross PCIe once).
 * results[n]; traces[trace_capacity] receives all traces back to back; *trace_used = entries
 * written (if it exceeds trace_capacity the call fails with GCGPU_ERR_ARG and *trace_used
 * tells the required size).  All buffers are host memory.                                  */
int gcgpu_extend(gcgpu_ctx* ctx, const uint8_t* seq, uint64_t seq_bytes, const gcgpu_ext_item* items, uint32_t n,
                 gcgpu_ext_result* results, uint64_t* traces, uint64_t trace_capacity, uint64_t* trace_used);
/* Two-phase form for callers that cannot bound the trace volume in advance: call gcgpu_extend with
 * traces == NULL and trace_capacity == 0 (results[] and *trace_used are filled, the traces stay on
 * the device), size a buffer, then copy entries [first, first+count) of the dense trace array of
 * that call.  Valid until the next gcgpu_extend on the same ctx.                              */
int gcgpu_fetch_traces(gcgpu_ctx* ctx, uint64_t* traces, uint64_t first, uint64_t count);


/* ---- S0: minimizer seeding lookups ------------------------------------------------------
 * Replaces, for a batch of reads, the k-mer walk and index probes of
 *   MinimizerSeeder::getSeeds -> iterateKmers -> addMinimizers
 * (src/MinimizerSeeder.cpp:522-544, 60-102, 494-520): for every read the list of
 * (k-mer END position, first index into the positions array, number of positions) of the
 * looked-up k-mers that are in the index with fewer than max_count positions, in ascending
 * position order -- the `matchIndices` vector before its sort by count.  The sort, the density cut
 * and matchToSeedHit (:533-555) stay with the caller (their std::sort tie order is libstdc++'s).
 * The index is the reference's: kmers[i] occurs at positions[kmer_start[i] .. kmer_start[i+1])
 * (the union over the MinimizerSeeder buckets; a k-mer lives in exactly one bucket).         */
typedef struct gcgpu_minimizer_index
{
	uint32_t k;                 /* minimizer length (15), <= 31            */
	uint32_t window;            /* window size (20); realWindow = window-k+1 */
	uint64_t max_count;         /* MinimizerSeeder::maxCount               */
	uint64_t num_kmers;
	const uint64_t* kmers;      /* [num_kmers]                             */
	const uint32_t* kmer_start; /* [num_kmers + 1]                         */
} gcgpu_minimizer_index;

typedef struct gcgpu_seed_read
{
	uint64_t seq_offset;        /* forward codes of the read inside `seq`  */
	int32_t seq_len;
	uint32_t reserved;
} gcgpu_seed_read;

typedef struct gcgpu_seed_match
{
	uint32_t pos;               /* k-mer END position in the read          */
	uint32_t start;             /* kmer_start[index]                       */
	uint32_t count;             /* kmer_start[index+1] - kmer_start[index] */
} gcgpu_seed_match;

int gcgpu_set_minimizer_index(gcgpu_ctx* ctx, const gcgpu_minimizer_index* index);
/* `seq` = the IUPAC-mask codes of gcgpu_extend, with bit 4 set on bases that are not one of
 * ACGTacgt for seeding although they match in the DP (U/u); the buffer stays resident for the
 * following gcgpu_extend calls (pass seq == NULL there).  match_offsets[n+1]; matches of read r are
 * [match_offsets[r], match_offsets[r+1]).  Two-phase use like gcgpu_extend: matches == NULL and
 * capacity == 0 leaves them on the device for gcgpu_fetch_seed_matches.                        */
int gcgpu_seed(gcgpu_ctx* ctx, const uint8_t* seq, uint64_t seq_bytes, const gcgpu_seed_read* reads, uint32_t n,
               uint64_t* match_offsets, gcgpu_seed_match* matches, uint64_t capacity, uint64_t* used);
int gcgpu_fetch_seed_matches(gcgpu_ctx* ctx, gcgpu_seed_match* matches, uint64_t first, uint64_t count);


/* ---- K3: global (NW) sequence alignment, Myers bit-vector ------------------------------
 * One item replaces one call of
 *   edlibAlign(query, qlen, target, tlen, edlibNewAlignConfig(-1, EDLIB_MODE_NW, task, NULL, 0))
 * (edlib/include/edlib.h, edlib/src/edlib.cpp:141-296) as made at src/Aligner.cpp:645
 * (TASK_DISTANCE, query = whole-read path) and src/Aligner.cpp:845 (TASK_PATH, query =
 * chained path, target = read).  `seqs` holds raw characters (compared byte-wise like edlib,
 * only upper-case A C G T can match).  want_path != 0 also returns the edit operations
 * 0 match / 1 insert / 2 delete / 3 mismatch (EDLIB_EDOP_*), identical to edlib's.         */
typedef struct gcgpu_nw_item
{
	uint64_t query_offset;
	uint64_t target_offset;
	int32_t query_len;
	int32_t target_len;
	int32_t k_hint;       /* optional first band guess (<= 0: start at 64 like edlib)      */
	int32_t want_path;    /* 0 distance only; 1 distance + edit operations; 2 edit operations for a pair whose
	                         exact edit distance the caller passes in k_hint (the value an earlier call returned):
	                         if every item of a call says 2 the distance pass is skipped                        */
} gcgpu_nw_item;

typedef struct gcgpu_nw_result
{
	int32_t status;       /* 0 ok, 5 internal */
	int32_t distance;     /* EdlibAlignResult::editDistance */
	uint32_t ops_len;     /* EdlibAlignResult::alignmentLength */
	uint32_t reserved;
	uint64_t ops_offset;  /* into the ops buffer */
	uint64_t blocks;      /* work counter: 64-row block column steps */
} gcgpu_nw_result;

/* `seqs` may be NULL: the buffer uploaded by the previous gcgpu_nw call on this ctx (same seq_bytes) is reused -- the edit
 * paths of a batch are asked for in a second call on the same pairs.                                              */
int gcgpu_nw(gcgpu_ctx* ctx, const char* seqs, uint64_t seq_bytes, const gcgpu_nw_item* items, uint32_t n,
             gcgpu_nw_result* results, uint8_t* ops, uint64_t ops_capacity, uint64_t* ops_used);

/* ---- K2: co-linear chaining over the minimum path cover ---------------------------------
 * One read replaces one call of
 *   AlignmentGraph::colinearChaining(const std::vector<Anchor>&, long long sep_limit)
 * (src/AlignmentGraph.h:121, src/AlignmentGraph.cpp:1712-1863; sep_limit is not read by the
 * reference).  An anchor is given by the split nodes of Anchor::path.front()/back() and its
 * fragment bounds x,y (src/Aligner.cpp:707).  Anchors of read r are
 * anchors[read_offsets[r] .. read_offsets[r+1]); the chain (anchor indices local to the read,
 * in read order) is written at chain[read_offsets[r] ..] with its length in chain_len[r]
 * and the covered-bases score in chain_score[r].                                            */
typedef struct gcgpu_anchor
{
	uint32_t start_node;
	uint32_t end_node;
	int32_t x;
	int32_t y;
} gcgpu_anchor;

int gcgpu_chain(gcgpu_ctx* ctx, const gcgpu_anchor* anchors, const uint64_t* read_offsets, uint32_t num_reads,
                uint32_t* chain, uint32_t* chain_len, int64_t* chain_score);

/* ---- resident batch: everything between the DP kernels, on the device --------------------
 * The entry points above return every K1 trace to the host (8 bytes per DP cell of every extension, ~66 bytes per read
 * base), where the reference's per-read logic then walks them.  The entry points below keep the traces of a batch in
 * HBM ("trace sets") and run that logic where the traces are; the host sees alignment extents, seed-coverage bit
 * masks, anchor records and edit runs (a few bytes per read base).
 *
 * gcgpu_load_reads        the reads' characters -> IUPAC codes (forward + reverse complement, what
 *                         CommonUtils::ReverseComplement + Common::ambiguousMatch give, GraphAlignerCommon.h:219-296)
 *                         and the byte codes K3 compares.  Codes of read r: forward at 2*char_offset, reverse
 *                         complement at 2*char_offset + len of the resident K1 sequence buffer (seq == NULL in
 *                         gcgpu_seed / gcgpu_extend with seq_bytes = 2 * char_bytes).
 * gcgpu_set_seed_cells    the batch's seeds (SeedHit, GraphAlignerWrapper.h:11-37) after OrderSeeds, per read in the
 *                         caller's order; reads[r].first_cell / num_cells delimit read r's cells.
 * gcgpu_extend_seeds      getAlignmentFromSeed (GraphAligner.h:567-626) for a list of seeds: both K1 extensions of
 *                         each seed, traces kept in trace set `set` (append != 0: after the pairs already there;
 *                         *first_pair = index of exts[0]'s pair in the set).  brief[i] = extent and score of the
 *                         merged alignment.  cover_bits != NULL: for every seed extension that produced an alignment,
 *                         bit c of the words at cover_word_offsets[i] says whether cell first_cell+c of the same read
 *                         lies on its trace (exactAlignmentPart, GraphAligner.h:407-461) -- the skip rule of
 *                         AlignOneWay (:162-173) then needs no trace on the host.
 * gcgpu_fragment_anchors  Aligner.cpp:668-730 for a batch: every window seed of every fragment extended (K1), the
 *                         in-order seed loop of AlignOneWay per fragment, anchors (Aligner.cpp:706-729) left on the
 *                         device for gcgpu_chain_resident.  frags must be grouped by read in ascending start.
 * gcgpu_chain_resident    gcgpu_chain on the anchors gcgpu_fragment_anchors left; gcgpu_fetch_chained returns the
 *                         chained anchors only (read order, chain order) with their node paths.
 * gcgpu_nw_compose        builds the resident K3 sequence buffer from pieces: a read, the padded graph path of a
 *                         whole-read alignment (traceToSequence, Aligner.cpp:376-408,425-428) or the bases of a
 *                         chained node path (pathToTrace, Aligner.cpp:409-424); gcgpu_nw(seqs == NULL) then runs on it.
 * gcgpu_encode_alignments GraphAlignerVGAlignment::traceToAlignment (GraphAlignerVGAlignment.h:37-165) for whole-read
 *                         alignments of a trace set: mappings and edit runs as a uint32 token stream
 *                         (mapping: 0, digraph node id, offset; edit: type << 30 | run length, type 0 match,
 *                         1 mismatch, 2 insertion, 3 deletion).                                                  */
#define GCGPU_TRACE_SETS 4

typedef struct gcgpu_read
{
	uint64_t char_offset;
	int32_t len;
	uint32_t first_cell;
	uint32_t num_cells;
	uint32_t reserved;
} gcgpu_read;

typedef struct gcgpu_seed_cell
{
	int32_t seq_pos;      /* SeedHit::seqPos (k-mer END position in the read)          */
	uint32_t node;        /* SeedHit::alignmentGraphNodeId (split node)                */
	uint32_t read;        /* index of the read in the batch                            */
	uint8_t offset;       /* SeedHit::alignmentGraphNodeOffset                         */
	uint8_t flags;        /* bit 0: seedClusterSize >= seedClusterMinSize              */
	uint16_t reserved;
} gcgpu_seed_cell;

typedef struct gcgpu_seed_ext
{
	uint32_t cell;        /* index into the batch's seed cells                         */
	int32_t frag_start;   /* first read position of the fragment; < 0: the whole read  */
} gcgpu_seed_ext;

#define GCGPU_PAIR_BWD 1u       /* the backward extension produced a trace               */
#define GCGPU_PAIR_FWD 2u       /* the forward extension produced a trace                */
#define GCGPU_PAIR_INTERNAL 4u  /* an extension hit a state the reference asserts on     */
typedef struct gcgpu_pair_brief
{
	int32_t start, end;   /* AlignmentItem::alignmentStart / alignmentEnd (coordinates of the aligned sequence) */
	int32_t score;        /* OnewayTrace::score                                        */
	uint32_t flags;       /* GCGPU_PAIR_*; neither BWD nor FWD: the extension failed   */
} gcgpu_pair_brief;

typedef struct gcgpu_frag
{
	uint32_t read;
	int32_t start;        /* l of Aligner.cpp:675                                      */
	uint32_t first_ext;   /* its window seeds exts[first_ext .. first_ext + num_exts)  */
	uint32_t num_exts;
} gcgpu_frag;

typedef struct gcgpu_read_anchors
{
	uint32_t anchors;            /* A.size()                                                          */
	uint32_t seeds_extended;     /* sum of alignments.seedsExtended over the fragments (Aligner.cpp:705) */
	uint32_t last_frag_extended; /* seedsExtended of the last fragment that ran                        */
	uint32_t dropped;            /* 1: a fragment hit an assertion-class state (cont = true, :695-703); the
	                                anchors of the fragments before it are kept                          */
} gcgpu_read_anchors;

typedef struct gcgpu_chained_anchor
{
	uint32_t first_offset;  /* Apos[i][0]: offset in the first path node                */
	uint32_t last_offset;   /* Apos[i][1]: offset in the last path node                 */
	uint64_t path_first;    /* Anchor::path = path_nodes[path_first .. path_first + path_len) */
	uint32_t path_len;
	uint32_t reserved;
} gcgpu_chained_anchor;

#define GCGPU_PIECE_READ 0u       /* index = read                                           */
#define GCGPU_PIECE_PAIR_PATH 1u  /* index = pair in trace set `set`                        */
#define GCGPU_PIECE_NODE_PATH 2u  /* path_nodes[first_node .. first_node + num_nodes), first / last offset */
typedef struct gcgpu_nw_piece
{
	uint32_t kind;
	uint32_t set;
	uint32_t index;
	uint32_t num_nodes;
	uint64_t first_node;
	uint32_t first_offset;
	uint32_t last_offset;
} gcgpu_nw_piece;

typedef struct gcgpu_aln_tokens
{
	uint64_t token_offset;
	uint32_t num_tokens;
	uint32_t matches;      /* identity = matches / steps (GraphAlignerVGAlignment.h:150)  */
	uint32_t steps;
	uint32_t reserved;
} gcgpu_aln_tokens;

/* GAM records on the device (writeGAMToQueue, src/Aligner.cpp:261-281): for every listed read one gzip member holding
 * varint count + {varint size, vg::Alignment}* of its listed whole-read alignments, in the given order -- the proto3 bytes the
 * reference's AddAlignment + replaceDigraphNodeIdsWithOriginalNodeIds + SerializeToString produce (sequence = the read's characters
 * as passed to gcgpu_load_reads).  gcgpu_set_node_names provides the GFA segment names of the original nodes (in the order of
 * gcgpu_graph.orig_ids).  member_offsets[n + 1]: the member of reads[i] is [member_offsets[i], member_offsets[i + 1]) of the
 * device buffer gcgpu_fetch_gam copies from; an empty range means this record has to be encoded by the caller. */
typedef struct gcgpu_gam_aln
{
	uint32_t pair;         /* seed extension in the trace set                         */
	int32_t start, end;    /* AlignmentItem::alignmentStart / alignmentEnd             */
	int32_t trace_score;   /* vg::Alignment::score                                      */
} gcgpu_gam_aln;
typedef struct gcgpu_gam_read
{
	uint32_t read;         /* index in the batch of gcgpu_load_reads                   */
	uint32_t first_aln, num_alns;
	uint32_t name_len;
	uint64_t name_offset;  /* the read's name inside `names`                           */
} gcgpu_gam_read;
int gcgpu_set_node_names(gcgpu_ctx* ctx, const uint32_t* name_offsets, const char* names);
int gcgpu_encode_gam(gcgpu_ctx* ctx, int set, const gcgpu_gam_read* reads, uint32_t n, const gcgpu_gam_aln* alns, uint32_t num_alns,
                     const char* names, uint64_t name_bytes, uint64_t* member_offsets, uint64_t* bytes_used);
int gcgpu_fetch_gam(gcgpu_ctx* ctx, uint8_t* out, uint64_t first, uint64_t count);

int gcgpu_load_reads(gcgpu_ctx* ctx, const char* chars, uint64_t char_bytes, const gcgpu_read* reads, uint32_t n);
int gcgpu_set_seed_cells(gcgpu_ctx* ctx, const gcgpu_seed_cell* cells, uint64_t num_cells, const gcgpu_read* reads, uint32_t n);
int gcgpu_extend_seeds(gcgpu_ctx* ctx, int set, int append, int32_t frag_len, const gcgpu_seed_ext* exts, uint32_t n, gcgpu_pair_brief* brief,
                       uint32_t* cover_bits, const uint64_t* cover_word_offsets, uint32_t* first_pair, uint64_t* columns);
int gcgpu_fragment_anchors(gcgpu_ctx* ctx, int set, int32_t frag_len, const gcgpu_seed_ext* exts, uint32_t num_exts, const gcgpu_frag* frags, uint32_t num_frags,
                           uint32_t num_reads, gcgpu_read_anchors* per_read, uint64_t* columns);
int gcgpu_chain_resident(gcgpu_ctx* ctx, uint32_t num_reads, uint32_t* chain_len, int64_t* chain_score, uint64_t* chained_total, uint64_t* path_nodes_total);
int gcgpu_fetch_chained(gcgpu_ctx* ctx, gcgpu_chained_anchor* anchors, uint32_t* path_nodes);
int gcgpu_nw_compose(gcgpu_ctx* ctx, const gcgpu_nw_piece* pieces, uint32_t n, const uint32_t* path_nodes, uint64_t num_path_nodes, uint64_t* piece_offsets);
int gcgpu_encode_alignments(gcgpu_ctx* ctx, int set, const uint32_t* pairs, uint32_t n, gcgpu_aln_tokens* out, uint64_t* tokens_used);
int gcgpu_fetch_tokens(gcgpu_ctx* ctx, uint32_t* tokens, uint64_t first, uint64_t count);
/* bytes copied host->device / device->host by this ctx so far */
void gcgpu_transfer_bytes(gcgpu_ctx* ctx, uint64_t* h2d, uint64_t* d2h);

/* Page-locked host memory for the caller-owned buffers above (cudaHostAlloc): copies to and from
 * such buffers run at full PCIe rate and asynchronously.  Plain malloc'd buffers work too. */
void* gcgpu_host_alloc(size_t bytes);
void gcgpu_host_free(void* p);

/* Measurement aid: sustained rate of independent 32-bit LOP3/IADD3 instructions on this device (thread-level
 * int32 ops per second), the pipe that bounds K1/K3 (SURVEY.md 8d asks for a measured integer peak).  */
int gcgpu_int_peak(gcgpu_ctx* ctx, double* int32_ops_per_s);

/* device time of the kernels of the last call on this ctx, in milliseconds (CUDA events) */
float gcgpu_last_kernel_ms(gcgpu_ctx* ctx);
/* number of kernel launches issued by this ctx so far */
uint64_t gcgpu_launch_count(gcgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
