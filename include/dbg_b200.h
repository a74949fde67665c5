/* dbg_b200.h — C ABI of the B200-native read -> unitig path (libdbg_b200.so).
 *
 * Drop-in boundary for ONE path of the crate `debruijn` (10XGenomics/rust-debruijn 0.3.4):
 *     filter::filter_kmers (CountFilter)  ->  compression::compress_kmers_with_hash (SimpleCompress)
 * The reference has no FFI; the boundary is two generic Rust functions plus three public data
 * layouts.  Every entry point below names the reference item it replaces (paths relative to the
 * crate root).  Plain pointers and sizes only; no exceptions or unwinding cross this ABI; every call
 * returns a dbg_status and leaves a message retrievable with dbg_last_error().
 *
 * Limits: k-mer / node indices are 32-bit inside one table or graph (fewer than 2^31 valid k-mers per GPU — per rank's shard in
 * the multi-GPU calls — and fewer than 2^31 nodes per graph; larger inputs return DBG_E_BADARG), K in [4, 64], counts saturate at
 * 65 535 as in the crate (src/filter.rs:57).
 *
 * Threading: calls block.  One dbg_ctx = one CUDA device + one stream; a ctx is used by one caller at
 * a time; several ctxs (one per GPU / per process) may coexist.  There is NO CPU fallback: without a
 * CUDA device dbg_ctx_create fails with DBG_E_CUDA.
 *
 * Data layouts (identical to the crate's):
 *   sequences   PackedDnaStringSet image (src/dna_string.rs:763-767, 383-399): `words` = 2-bit bases,
 *               32 per u64, base b of the concatenation at bits 62-2(b%32) of word b/32;
 *               `start[i]` base offset, `length[i]` bases; `seq_exts[i]` = Exts.val of sequence i
 *               (NULL => Exts::empty()).
 *   k-mers      K<=32: u64, right-aligned, base 0 most significant (src/kmer.rs:429-437);
 *               32<K<=64: {lo,hi} = Rust u128 little-endian halves.
 *   Exts        u8, bits 0..3 left A,C,G,T, bits 4..7 right A,C,G,T (src/lib.rs:569-580).
 *   BaseGraph   sequences (bit-contiguous, no per-node padding), start (base offset), length (u32),
 *               exts (u8), data (u16), stranded (src/graph.rs:44-50).
 */
#ifndef DBG_B200_H
#define DBG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DBG_OK = 0,
    DBG_E_BADARG = 1,
    DBG_E_OOM = 2,
    DBG_E_CUDA = 3,
    DBG_E_INCONSISTENT_EXTS = 4, /* the reference's panic!("unreachable"), src/compression.rs:428-434 */
    DBG_E_INTERNAL = 5
} dbg_status;

/* SimpleCompress reduce closures used by the reference's tests (src/compression.rs:40-65). */
typedef enum {
    DBG_REDUCE_SAT_ADD = 0,       /* |a,b| a.saturating_add(*b)              src/test.rs:383,459 */
    DBG_REDUCE_WRAP_ADD = 1,      /* |a,b| a + b (release: wrapping)          src/test.rs:265,546 */
    DBG_REDUCE_ADD_MOD_65535 = 2, /* |a,b| ((a as u32 + *b as u32) % 65535)   src/test.rs:247     */
    DBG_REDUCE_MAX = 3,           /* |a,b| max(a,*b)                          src/test.rs:469     */
    DBG_REDUCE_SCMAP = 4          /* ScmapCompress (src/compression.rs:66-98): join_test = data equality, reduce keeps the value;
                                     also honoured by dbg_reads_to_graph_multi: the join_test crosses ranks) */
} dbg_reduce_op;

typedef struct dbg_ctx dbg_ctx;
typedef struct dbg_seqset dbg_seqset;      /* device-resident &[(V, Exts, D1)]            */
typedef struct dbg_kmer_table dbg_kmer_table; /* device-resident BoomHashMap2<K,Exts,u16> (+ all_kmers) */
typedef struct dbg_graph dbg_graph;        /* device-resident BaseGraph<K,u16>             */

/* Counters and per-stage device times (ms, CUDA events) of the last filter/compress call. */
typedef struct {
    uint64_t n_seqs, n_input_kmers, n_records, n_buckets, n_distinct, n_valid, n_nodes, n_bases;
    uint64_t n_bucket_splits; /* shared-memory table overflows resolved by hash-class splitting */
    uint64_t rank_rounds;     /* pointer-doubling rounds in compress */
    uint64_t n_cycle_kmers;   /* k-mers on cyclic unitigs */
    uint64_t gpu_launches;    /* kernels launched by this ctx so far */
    float ms_partition, ms_count, ms_sort, ms_table, ms_links, ms_rank, ms_emit;
    uint32_t msp_p, bucket_bits;
    float ms_k_partition, ms_k_count; /* the two dominant kernels alone (events right around the launch) */
    float ms_filter_total, ms_compress_total;
    uint64_t n_records_distinct; /* super-k-mer records left after per-bucket deduplication (0 = dedup off) */
    uint64_t n_passes;           /* passes over the reads chosen by the memory planner (filter.rs:151-168) */
    uint64_t direct_partition;   /* 1 = the last filter call wrote records straight into per-bucket regions (no staging / scatter) */
} dbg_stats;

/* ---- context ------------------------------------------------------------------------------------ */
int dbg_ctx_create(int device, dbg_ctx** out);
void dbg_ctx_destroy(dbg_ctx* ctx);
const char* dbg_last_error(const dbg_ctx* ctx);
int dbg_stats_get(const dbg_ctx* ctx, dbg_stats* out);
/* tunables: "msp_p" (minimizer length, 0 = auto), "bucket_occ" (target k-mer occurrences per MSP
 * bucket, 0 = auto), "dedup" (0 = off, 1 = 16-byte records [default], 2 = also 32-byte records), "mem_budget_bytes" (scratch budget of the pass planner, 0 = auto), "direct_partition" (1 = default: large contiguous inputs are
 * partitioned straight into per-bucket regions sized by a sampling pass; 0 = always stage + scatter), "direct_min_tiles"
 * (smallest input, in 4096-base tiles, that takes the direct partition; default 2048), "fast_compress" (1 = default; 0 = always use the general per-k-mer rank + emit
 * path of compress_kmers, which otherwise only runs when long unitigs or cycles are present). */
int dbg_ctx_set_param(dbg_ctx* ctx, const char* name, int64_t value);
int dbg_ctx_synchronize(dbg_ctx* ctx);
/* The cudaStream_t every call of this ctx is ordered on (for event timing / interop with other runtimes). */
void* dbg_ctx_stream(dbg_ctx* ctx);

/* ---- sequences: the `seqs: &[(V, Exts, D1)]` argument of filter_kmers (src/filter.rs:139-140) ----
 * D1 is not transported: CountFilter never reads it (src/filter.rs:56).  Host pointers are borrowed
 * for the duration of the call only. */
int dbg_seqset_upload(dbg_ctx* ctx, const uint64_t* words, uint64_t n_words, const uint64_t* start,
                      const uint32_t* length, const uint8_t* seq_exts, uint64_t n_seqs, dbg_seqset** out);
/* Fixed-length reads packed back to back (sequence i = bases [i*read_len, (i+1)*read_len)): the common
 * sequencer layout; no start/length arrays to transfer or validate. */
int dbg_seqset_upload_uniform(dbg_ctx* ctx, const uint64_t* words, uint64_t n_words, uint64_t n_seqs, uint32_t read_len,
                              const uint8_t* seq_exts, dbg_seqset** out);
/* Ingest — DnaString::from_acgt_bytes (src/dna_string.rs:224-250; AVX2 twin src/bitops_avx2.rs:8-132) for a batch of ASCII
 * sequences, appended like PackedDnaStringSet::add (src/dna_string.rs:811-821): sequence i = ascii[start[i] .. +length[i])
 * (host buffer, byte offsets); A/a C/c G/g T/t -> 0..3, anything else -> A; *n_invalid (optional) counts those.  The 2-bit
 * packing runs on the device; the result is a contiguous sequence set ready for dbg_filter_kmers. */
int dbg_seqset_from_ascii(dbg_ctx* ctx, const uint8_t* ascii, uint64_t n_bytes, const uint64_t* start, const uint32_t* length,
                          const uint8_t* seq_exts, uint64_t n_seqs, uint64_t* n_invalid, dbg_seqset** out);
/* DnaString::from_acgt_bytes_hashn (src/dna_string.rs:254-278): as above, but a non-ACGT character at position pos of a sequence
 * becomes the repeatable pseudo-random base  DefaultHasher(read_name, pos) % 4  — SipHash-1-3 with a zero key over
 * [len(name) u64 LE][name][pos u64 LE], the byte stream Rust's Hash impls for [u8] and usize feed the hasher.  names: all read
 * names back to back (host), name i = names[name_start[i] .. + name_len[i]). */
int dbg_seqset_from_ascii_hashn(dbg_ctx* ctx, const uint8_t* ascii, uint64_t n_bytes, const uint64_t* start, const uint32_t* length,
                                const uint8_t* names, uint64_t n_name_bytes, const uint64_t* name_start, const uint32_t* name_len,
                                const uint8_t* seq_exts, uint64_t n_seqs, uint64_t* n_invalid, dbg_seqset** out);
/* SipHash-1-3, zero key, of a byte string (host): the primitive behind the call above, exported so that it can be pinned against
 * an independent implementation. */
uint64_t dbg_siphash13(const uint8_t* bytes, uint64_t n);
/* Same, asynchronous: returns at once, the packed words go up in chunks on a copy stream and the partition stage of
 * the next dbg_filter_kmers / dbg_partition_reads / dbg_reads_to_graph call starts on the chunks that have arrived.
 * `words` (pinned host memory for real overlap) must stay valid and unchanged until that call has returned. */
int dbg_seqset_upload_uniform_async(dbg_ctx* ctx, const uint64_t* words, uint64_t n_words, uint64_t n_seqs, uint32_t read_len,
                                    const uint8_t* seq_exts, dbg_seqset** out);
/* Wrap caller-owned DEVICE buffers (e.g. torch tensors) without copying. */
int dbg_seqset_wrap_device(dbg_ctx* ctx, const uint64_t* d_words, uint64_t n_words, const uint64_t* d_start,
                           const uint32_t* d_length, const uint8_t* d_seq_exts, uint64_t n_seqs, uint32_t max_len,
                           dbg_seqset** out);
/* synth-v1 generator on the device: R reads x 150 bases. */
int dbg_seqset_synth(dbg_ctx* ctx, uint64_t n_reads, uint64_t seed, uint32_t err_thr, dbg_seqset** out);
uint64_t dbg_seqset_len(const dbg_seqset* s);
uint64_t dbg_seqset_n_words(const dbg_seqset* s);
int dbg_seqset_copy_out(const dbg_seqset* s, uint64_t* words, uint64_t* start, uint32_t* length);
void dbg_seqset_free(dbg_seqset* s);

/* ---- filter::filter_kmers with CountFilter::new(min_kmer_obs) — src/filter.rs:139-231, 40-63 -----
 * Result = the (valid_kmers, valid_exts, valid_data) arrays in ascending k-mer order, i.e. exactly
 * what the reference hands to BoomHashMap2::new (src/filter.rs:227-230), plus `all_kmers` when
 * report_all_kmers != 0.  memory_size_gb mirrors the reference argument; it never changes results
 * (src/filter.rs:151-168). */
int dbg_filter_kmers(dbg_ctx* ctx, int k, const dbg_seqset* seqs, uint32_t min_kmer_obs, int stranded,
                     int report_all_kmers, uint64_t memory_size_gb, dbg_kmer_table** out);
/* Same, host buffers in (upload + filter in one call; the e2e entry point). */
int dbg_filter_kmers_host(dbg_ctx* ctx, int k, const uint64_t* words, uint64_t n_words, const uint64_t* start,
                          const uint32_t* length, const uint8_t* seq_exts, uint64_t n_seqs, uint32_t min_kmer_obs,
                          int stranded, int report_all_kmers, uint64_t memory_size_gb, dbg_kmer_table** out);
/* filter::filter_kmers with CountFilterSet<u8>::new(min_kmer_obs) — src/filter.rs:68-101: `labels` (host, one per sequence, < 64) is
 * the D1 of the reference's (V, Exts, D1) tuples; the summary of a k-mer is the sorted, deduplicated set of the labels it was
 * observed with, returned as a 64-bit mask per k-mer (dbg_table_colorsets: bit c = label c) — the content of the reference's
 * Vec<u8> after sort() + dedup().  Valid iff observations >= min_kmer_obs (<= 65535).  Exts as for CountFilter; the table's
 * counts hold min(observations, 65535). */
int dbg_filter_kmers_colorset(dbg_ctx* ctx, int k, const dbg_seqset* seqs, const uint8_t* labels, uint32_t min_kmer_obs, int stranded,
                              uint64_t memory_size_gb, dbg_kmer_table** out);
int dbg_table_colorsets(const dbg_kmer_table* t, uint64_t* masks /* dbg_table_len entries */);
uint64_t dbg_table_len(const dbg_kmer_table* t);      /* BoomHashMap2::len */
uint64_t dbg_table_all_len(const dbg_kmer_table* t);  /* all_kmers.len()   */
uint64_t dbg_table_n_input(const dbg_kmer_table* t);
int dbg_table_k(const dbg_kmer_table* t);
/* Any pointer may be NULL.  kmers_hi / all_hi are only written for k > 32. */
int dbg_table_copy_out(const dbg_kmer_table* t, uint64_t* kmers_lo, uint64_t* kmers_hi, uint8_t* exts,
                       uint16_t* counts, uint64_t* all_lo, uint64_t* all_hi);
/* The `kmer_exts: &[(K,(Exts,D))]` argument of compression::compress_kmers (src/compression.rs:598-615):
 * any order, distinct k-mers; sorted ascending on the device (seed order). */
int dbg_table_from_host(dbg_ctx* ctx, int k, uint64_t n, const uint64_t* kmers_lo, const uint64_t* kmers_hi,
                        const uint8_t* exts, const uint16_t* counts, dbg_kmer_table** out);
void dbg_table_free(dbg_kmer_table* t);

/* ---- compression::compress_kmers_with_hash with SimpleCompress — src/compression.rs:588-594 -------
 * Node order = ascending smallest k-mer of each unitig (seed order = ascending canonical k-mer). */
int dbg_compress_kmers_with_hash(dbg_ctx* ctx, const dbg_kmer_table* index, int stranded, int reduce_op,
                                 dbg_graph** out);
uint64_t dbg_graph_len(const dbg_graph* g);      /* BaseGraph::len, src/graph.rs:63-65 */
uint64_t dbg_graph_n_bases(const dbg_graph* g);  /* sequences.sequence.len()           */
uint64_t dbg_graph_n_words(const dbg_graph* g);  /* ceil(n_bases / 32)                  */
int dbg_graph_stranded(const dbg_graph* g);
int dbg_graph_k(const dbg_graph* g);
int dbg_graph_copy_out(const dbg_graph* g, uint64_t* words, uint64_t* start, uint32_t* length, uint8_t* exts,
                       uint16_t* data);
/* serde image of BaseGraph<K, u16> in bincode 1.x's default encoding (little-endian fixed-width integers, u64 sequence lengths),
 * field order as derived in the crate (src/graph.rs:43-50; src/dna_string.rs:72-76, 762-767; src/lib.rs:577-580):
 *   sequences.sequence.storage | sequences.sequence.len | sequences.start | sequences.length | exts | data | stranded
 * i.e. what `bincode::serialize(&base_graph)` writes and `bincode::deserialize::<BaseGraph<K, u16>>` reads.  Two-phase:
 * *n_bytes is always set, `out` is written only when cap >= *n_bytes.  K is a type parameter in the crate: deserialize takes k. */
int dbg_graph_serialize(const dbg_graph* g, uint8_t* out, uint64_t cap, uint64_t* n_bytes);
int dbg_graph_deserialize(dbg_ctx* ctx, int k, const uint8_t* bytes, uint64_t n_bytes, dbg_graph** out);
/* BaseGraph::finish + DebruijnGraph::find_edges for EVERY (node, side) — src/graph.rs:116-142 (left_order / right_order),
 * :223-291 (find_edges, find_link).  Host outputs of 8 * n_nodes entries, slot (node * 2 + side) * 4 + base (side 0 = Left,
 * 1 = Right; base = A, C, G, T): target = node the extension leads to (0xffffffff: the node has no such extension, or the
 * link is not in this graph — "this edge doesn't exist within this shard", graph.rs:236), flags bit 0 = side of the target
 * through which it is entered (0 Left, 1 Right), bit 1 = reverse-complement switch. */
int dbg_graph_edges(dbg_ctx* ctx, const dbg_graph* graph, uint32_t* target, uint8_t* flags);
/* DebruijnGraph::fix_exts / get_valid_exts — src/graph.rs:337-377: rewrites the graph's node Exts in place, keeping an
 * extension only if find_link resolves it (and, with valid_nodes != NULL — host, one byte per node, the reference's BitSet —
 * only if the target node is marked).  "Remove non-existent extensions that may be created due to filtered kmers". */
int dbg_graph_fix_exts(dbg_ctx* ctx, dbg_graph* graph, const uint8_t* valid_nodes);
/* DebruijnGraph::is_compressed — src/graph.rs:296-334, the property the reference's tests assert after compress_kmers
 * (src/test.rs:248-254, 266-274): *pair_out = -1 when no two nodes could be collapsed, else (node << 32) | next_node for the FIRST
 * pair in the reference's iteration order.  scmap_join_test != 0: spec.join_test = data equality (ScmapCompress), else always true. */
int dbg_graph_is_compressed(dbg_ctx* ctx, const dbg_graph* graph, int scmap_join_test, int64_t* pair_out);
/* BaseGraph::combine — src/graph.rs:71-100: the nodes of every graph re-added in order (bases re-packed contiguously, Exts and data
 * concatenated).  Graphs of one K; mixing stranded and unstranded graphs (a panic there, :90-92) returns DBG_E_BADARG. */
int dbg_graph_combine(dbg_ctx* ctx, const dbg_graph* const* graphs, uint32_t n_graphs, dbg_graph** out);
/* compression::compress_graph — src/compression.rs:291-349 (CompressFromGraph :100-287): fix_exts(Some(available)) on a copy of the
 * Exts, merge every unbranched run of available nodes (censor_nodes: host array of node ids that are not available, may be NULL),
 * finish + fix_exts(None).  `stranded` is the function's own argument (palindrome rule, flag of the result); links are resolved under
 * the input graph's flag, as DebruijnGraph::find_link does.  Output nodes in the greedy loop's order: by the smallest node id of the
 * run, that node forward, a cycle of nodes opened so that it ends on it.  reduce_op as for dbg_compress_kmers_with_hash.  A graph whose node links are not reciprocal (the
 * greedy result would depend on the walk order) or that trips the reference's "unreachable" panic returns DBG_E_INCONSISTENT_EXTS. */
int dbg_compress_graph(dbg_ctx* ctx, const dbg_graph* graph, int stranded, int reduce_op, const uint64_t* censor_nodes,
                       uint64_t n_censor, dbg_graph** out);
void dbg_graph_free(dbg_graph* g);

/* ---- fused path: reads -> BaseGraph with the k-mer table kept device-resident ----------------------
 * Equivalent to filter_kmers(...) followed by compress_kmers_with_hash(...) (src/test.rs:344-386).
 * `table_out` may be NULL. */
int dbg_reads_to_graph(dbg_ctx* ctx, int k, const dbg_seqset* seqs, uint32_t min_kmer_obs, int stranded,
                       int reduce_op, dbg_kmer_table** table_out, dbg_graph** graph_out);
/* start == NULL && length == NULL: fixed-length reads of `uniform_len` bases (see dbg_seqset_upload_uniform). */
int dbg_reads_to_graph_host_uniform(dbg_ctx* ctx, int k, const uint64_t* words, uint64_t n_words, uint64_t n_seqs,
                                    uint32_t read_len, const uint8_t* seq_exts, uint32_t min_kmer_obs, int stranded,
                                    int reduce_op, dbg_kmer_table** table_out, dbg_graph** graph_out);
int dbg_reads_to_graph_host(dbg_ctx* ctx, int k, const uint64_t* words, uint64_t n_words, const uint64_t* start,
                            const uint32_t* length, const uint8_t* seq_exts, uint64_t n_seqs, uint32_t min_kmer_obs,
                            int stranded, int reduce_op, dbg_kmer_table** table_out, dbg_graph** graph_out);

/* ---- building blocks of the bucket-sharded counting stage (the reference's sharded flow, src/test.rs:433-456: msp_sequence
 * -> per-shard filter_kmers), exported for callers that run their own exchange: every rank partitions its own reads with the
 * same plan, ships each bucket range to its owner and counts the buckets it owns.  dbg_reads_to_graph_multi (above) does all
 * of this inside the library. */
typedef struct dbg_partition dbg_partition;
/* Minimizer length and log2(#buckets) for a job of n_input_kmers_total k-mer occurrences (same on every rank). */
int dbg_plan_filter(dbg_ctx* ctx, int k, uint64_t n_input_kmers_total, int* msp_p, int* bucket_bits);
uint64_t dbg_seqset_count_kmers(dbg_ctx* ctx, int k, const dbg_seqset* seqs);
int dbg_partition_reads(dbg_ctx* ctx, int k, const dbg_seqset* seqs, int stranded, int msp_p, int bucket_bits,
                        dbg_partition** out);
uint64_t dbg_partition_n_records(const dbg_partition* p);
uint64_t dbg_partition_n_input(const dbg_partition* p);
uint32_t dbg_partition_record_bytes(const dbg_partition* p);       /* 16 (k <= 32) or 32 */
void* dbg_partition_records_dev(const dbg_partition* p);          /* device pointer, records in bucket order */
int dbg_partition_bucket_counts(const dbg_partition* p, uint32_t* host_counts /* 2^bucket_bits */);
void dbg_partition_free(dbg_partition* p);
/* d_records (device): n_src runs back to back; run s holds this rank's n_local_buckets buckets in bucket order;
 * h_counts[s * n_local_buckets + b] (host) = records of local bucket b in run s. */
int dbg_filter_from_records(dbg_ctx* ctx, int k, const void* d_records, uint64_t n_records, const uint32_t* h_counts,
                            uint32_t n_src, uint32_t n_local_buckets, uint64_t n_input_kmers_total,
                            uint32_t min_kmer_obs, int stranded, int report_all_kmers, dbg_kmer_table** out);
/* A table of n entries with UNINITIALISED device arrays: the caller fills them (ascending, distinct k-mers) through
 * dbg_table_device_ptrs, e.g. as the receive buffers of a collective. */
int dbg_table_alloc(dbg_ctx* ctx, int k, uint64_t n, dbg_kmer_table** out);
/* filter::remove_censored_exts (src/filter.rs:280-306; sharded = 0) and remove_censored_exts_sharded (:238-276; sharded = 1):
 * rewrites the table's Exts in place, dropping every extension that points at a k-mer which is not in the table (plain), or
 * which is not in the table but is in the table's all_kmers (sharded; needs report_all_kmers at filter time). */
int dbg_remove_censored_exts(dbg_ctx* ctx, dbg_kmer_table* table, int stranded, int sharded);
/* Histogram of the top `bits` (<= min(24, 2k)) bits of the table's keys into d_hist (device, 2^bits u32, zeroed by the
 * call; stream-ordered, not synchronised): splitters for redistributing sorted shards by key range. */
int dbg_table_prefix_hist(dbg_ctx* ctx, const dbg_kmer_table* t, int bits, void* d_hist);
/* Device pointers of a table's arrays (hi is NULL for k <= 32), and a table built from device arrays (any order). */
int dbg_table_device_ptrs(const dbg_kmer_table* t, void** kmers_lo, void** kmers_hi, void** exts, void** counts);
int dbg_table_from_device(dbg_ctx* ctx, int k, uint64_t n, const void* d_kmers_lo, const void* d_kmers_hi,
                          const void* d_exts, const void* d_counts, dbg_kmer_table** out);
/* Same, for arrays that are ALREADY ascending and distinct (verified on the device, DBG_E_BADARG otherwise): no sort. */
int dbg_table_from_device_sorted(dbg_ctx* ctx, int k, uint64_t n, const void* d_kmers_lo, const void* d_kmers_hi,
                                 const void* d_exts, const void* d_counts, dbg_kmer_table** out);

/* A BaseGraph handle from caller-owned device arrays (copied). */
int dbg_graph_from_device(dbg_ctx* ctx, int k, int stranded, uint64_t n_nodes, uint64_t n_bases, const void* d_words,
                          const void* d_start, const void* d_length, const void* d_exts, const void* d_data,
                          dbg_graph** out);

/* ---- multi-GPU entry points: the whole path over several GPUs INSIDE the library (NCCL over NVLink, CUDA IPC peer windows) ----
 * Replaces the per-shard flow of the reference's sharded test (src/test.rs:418-470: msp_sequence -> per-shard filter_kmers ->
 * per-shard compress -> BaseGraph::combine + compress_graph) with one collective call whose per-rank outputs, concatenated in
 * rank order, are the single-GPU BaseGraph bit for bit (node order = ascending smallest k-mer).
 * A communicator is one rank: one process per GPU (dbg_comm_create; every rank passes the same 128-byte id obtained from
 * dbg_comm_unique_id on one rank and distributed by the caller's own rendezvous), or one process driving several GPUs with one
 * host thread per rank (dbg_multi_*; the communicators are owned by the handle).  Calls are collective: every rank must make
 * the same call.  Transport "nccl" = peer windows (CUDA IPC) written directly by the exchange kernel over NVLink, grouped
 * ncclSend/ncclRecv for the small exchanges, all-gather / all-reduce on the ctx stream; "local" (dbg_multi_* only, when
 * the device list repeats a device, NCCL is missing, or DBG_MULTI_TRANSPORT=local) stages through cudaMemcpy between the ranks'
 * buffers and exists so that the multi-rank logic can be exercised on ONE GPU. */
typedef struct dbg_comm dbg_comm;
typedef struct dbg_multi dbg_multi;
typedef struct {
    uint32_t n_ranks, rank;
    uint64_t n_input_total;  /* k-mer occurrences of the whole job */
    uint64_t n_valid_total, n_valid_local;   /* valid k-mers: whole job / this rank's shard of the table */
    uint64_t n_nodes_total, n_bases_total;   /* the complete BaseGraph */
    uint64_t node0, base0;   /* position of this rank's run of nodes in the complete graph (0 when replicated) */
    uint64_t n_queries_sent; /* neighbour lookups this rank had to send to other ranks */
    uint64_t exchange_bytes_sent; /* super-k-mer record bytes this rank stored into other ranks' windows (the path's one big exchange) */
    uint32_t replicated;     /* 1 = long unitigs / cycles: the table was gathered and every rank returns the COMPLETE graph */
    uint32_t check_ok;       /* n_bases_total == n_valid_total + n_nodes_total * (k - 1) and every k-mer was covered */
    uint32_t msp_p, bucket_bits;
    float ms_partition, ms_exchange, ms_count_sort, ms_links, ms_discover, ms_layout, ms_emit, ms_total;
} dbg_multi_info;
int dbg_comm_unique_id(void* id_out /* 128 bytes */);
int dbg_comm_create(dbg_ctx* ctx, int n_ranks, int rank, const void* unique_id, dbg_comm** out);
void dbg_comm_destroy(dbg_comm* comm);
int dbg_comm_rank(const dbg_comm* comm);
int dbg_comm_size(const dbg_comm* comm);
const char* dbg_comm_transport(const dbg_comm* comm);
/* filter_kmers(CountFilter) + compress_kmers_with_hash(SimpleCompress / ScmapCompress) over ALL ranks' sequences.  graph_out = this
 * rank's run of nodes (start[] relative to the run; info->node0 / base0 place it), or the complete graph when info->replicated. */
int dbg_reads_to_graph_multi(dbg_comm* comm, int k, const dbg_seqset* seqs, uint32_t min_kmer_obs, int stranded, int reduce_op,
                             dbg_multi_info* info, dbg_graph** graph_out);
/* one process, n ranks: rank i runs on devices[i] with its own ctx (dbg_multi_ctx: create that rank's sequence set on it) */
int dbg_multi_create(const int* devices, int n, dbg_multi** out);
void dbg_multi_destroy(dbg_multi* m);
int dbg_multi_size(const dbg_multi* m);
dbg_ctx* dbg_multi_ctx(dbg_multi* m, int rank);
const char* dbg_multi_transport(const dbg_multi* m);
int dbg_multi_reads_to_graph(dbg_multi* m, int k, const dbg_seqset* const* seqs, uint32_t min_kmer_obs, int stranded, int reduce_op,
                             dbg_multi_info* infos /* n or NULL */, dbg_graph** graphs_out /* n */);
/* pure host helpers of the plan (no GPU): bucket ownership (rank r owns [bounds[r], bounds[r+1])) and quantile cuts of a histogram */
int dbg_plan_owner_bounds(uint64_t n_buckets, int n_ranks, uint64_t* bounds_out /* n_ranks + 1 */);
int dbg_plan_quantile_cuts(const uint64_t* hist, uint64_t n_bins, int n_ranks, uint64_t* cuts_out /* n_ranks + 1 */);
/* Layout of the fused record exchange of dbg_reads_to_graph_multi, as a pure host function (the device computes the same from the
 * all-gathered counts): all_counts[s * n_buckets + b] = records of bucket b on rank s; dst_off[b] = where sender `me` stores its
 * records of bucket b inside the window of the bucket's owner (in records; buckets contiguous, senders in rank order inside a
 * bucket); recv_total[r] = records rank r ends up with. */
int dbg_plan_exchange_layout(const uint32_t* all_counts, int n_ranks, int me, uint64_t n_buckets, uint64_t* dst_off /* n_buckets */,
                             uint64_t* recv_total /* n_ranks */);

/* ---- msp::msp_sequence bucket assignment — src/msp.rs:279-324, 115-117 -----------------------------
 * For every k-mer start position j of every sequence: the MSP bucket of that k-mer under the
 * reference's default (identity) permutation with rc = !stranded, i.e. the value
 * MspIntervalP::bucket() of the interval that Scanner::scan places the k-mer in.
 * out_bucket has sum_i max(length[i]-k+1, 0) entries, sequence-major. */
int dbg_msp_kmer_buckets(dbg_ctx* ctx, int k, int p, const dbg_seqset* seqs, int stranded, uint32_t* out_bucket,
                         uint64_t n_out);

/* msp::msp_sequence — src/msp.rs:279-324 with Scanner::scan (:207-276): the MSP intervals of every sequence, in scan order, as
 * (sequence index, start, len [the Vmer the reference copies out is bases start .. start+len of that sequence], bucket =
 * MspIntervalP::bucket() as u32, Exts::from_slice_bounds src/lib.rs:645-660).  permutation: host array of 4^p scores (NULL = the
 * reference's default identity permutation; tables up to p = 12), rc = the reference's `rc` argument.  Sequences shorter than k
 * yield nothing.  Two-phase: *n_intervals is always set; the arrays are written only when cap >= *n_intervals. */
int dbg_msp_sequence(dbg_ctx* ctx, int k, int p, const dbg_seqset* seqs, int rc, const uint32_t* permutation, uint64_t cap,
                     uint64_t* n_intervals, uint32_t* seq, uint32_t* start, uint32_t* len, uint32_t* bucket, uint8_t* exts);

#ifdef __cplusplus
}
#endif
#endif /* DBG_B200_H */
