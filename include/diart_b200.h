/*
 * diart_b200 -- C ABI of the B200-native diart hot path (libdiartb200.so, sm_100a).
 *
 * This is the drop-in boundary.  The reference is pure Python and has
 * no FFI of its own; each entry point below names the reference interface it replaces and is
 * what a ctypes/cffi binding of that interface would call.
 *
 * Conventions
 *   - every function returns 0 on success, a negative DG_E* code on failure; the message is
 *     available from dg_last_error() (thread-local).
 *   - pointers documented "dev" are device pointers on the handle's device; "host" are host
 *     pointers.  Device entry points are stream-ordered on `stream` (a cudaStream_t passed as
 *     void*, NULL = legacy default stream) and never synchronise, except where noted.
 *   - a handle is not thread-safe; distinct handles are independent.  This matches the
 *     reference, where a pipeline instance is only ever driven by one thread
 *     (reference src/diart/inference.py:230).
 *   - tensors are dense, row-major, float32 unless stated otherwise.
 */
#ifndef DIART_B200_H
#define DIART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_OK 0
#define DG_EINVAL (-1)   /* bad argument / shape (the reference raises AssertionError / ValueError) */
#define DG_ECUDA (-2)    /* CUDA runtime error */
#define DG_EWEIGHT (-3)  /* missing / mis-shaped tensor in the state dict */
#define DG_ENCCL (-4)

/* One named float32 tensor of a pyannote state_dict (host memory), e.g.
 * {"sincnet.conv1d.1.weight", ptr, 24000}.  Key names follow pyannote.audio's PyanNet /
 * XVectorSincNet modules -- what reference src/diart/models.py:50 loads. */
typedef struct dg_tensor {
  const char* name;
  const float* data;
  int64_t numel;
} dg_tensor;

typedef struct dg_seg dg_seg;
typedef struct dg_emb dg_emb;
typedef struct dg_cluster dg_cluster;
typedef struct dg_pipeline dg_pipeline;
typedef struct dg_post dg_post;
typedef struct dg_stream dg_stream;

const char* dg_last_error(void);
int dg_version(void);

/* ---- segmentation model: replaces the callable behind SegmentationModel.__call__
 *      (reference src/diart/models.py:188-198; call site src/diart/blocks/segmentation.py:47):
 *      waveform (B,1,S) -> (B,F,K) sigmoid scores. ---- */
int dg_seg_create(const dg_tensor* tensors, int n_tensors, int device, dg_seg** out);
/* frames / local speakers produced for `num_samples`-sample chunks (293 / 3 for 80000) */
int dg_seg_dims(const dg_seg* h, int num_samples, int* frames, int* speakers);
/* Powerset models (pyannote/segmentation-3.0; reference PowersetAdapter, src/diart/models.py:29-39): the classifier has
 * one output per subset of the `num_speakers` local speakers of size <= `max_per_frame` (pyannote Powerset order: by
 * size, then lexicographic); forward then returns the hard multilabel scores one_hot(argmax) @ mapping, (B, F,
 * num_speakers), and dg_seg_dims reports num_speakers. */
int dg_seg_set_powerset(dg_seg* h, int num_speakers, int max_per_frame);
int dg_seg_forward(dg_seg* h, const float* wav_dev /*[B,S]*/, int B, int S,
                   float* seg_dev /*[B,F,K]*/, void* stream);
int dg_seg_destroy(dg_seg* h);

/* ---- embedding model: replaces the callable behind EmbeddingModel.__call__
 *      (reference src/diart/models.py:248-265; call site src/diart/blocks/embedding.py:60-67).
 *      pool_mode: 31 = pyannote.audio 3.1 StatsPool (nearest resize, +1e-8), 21 = 2.1 (linear). ---- */
int dg_emb_create(const dg_tensor* tensors, int n_tensors, int pool_mode, int device, dg_emb** out);
int dg_emb_dims(const dg_emb* h, int num_samples, int* frames, int* dimension);
/* Fused form: one trunk pass per waveform, K weighted poolings.
 * weights_dev [B,F,K] (the layout OverlappedSpeechPenalty returns) or NULL (unweighted, K must be 1).
 * If normalize != 0 rows are L2-normalised to `norm` (EmbeddingNormalization,
 * reference src/diart/blocks/embedding.py:110-120).  out_dev [B,K,D]. */
int dg_emb_forward(dg_emb* h, const float* wav_dev /*[B,S]*/, const float* weights_dev, int B, int S,
                   int F, int K, int normalize, float norm, float* out_dev, void* stream);
/* Compatibility form, exactly the arguments the reference block passes
 * (src/diart/blocks/embedding.py:57-65): waveform rows already repeated K times, weights (N,F).
 * Consecutive identical waveform rows are detected on the device and share one trunk pass.
 * Performs one small D2H read (row-group flags), i.e. synchronises `stream`. */
int dg_emb_forward_rows(dg_emb* h, const float* wav_dev /*[N,S]*/, const float* weights_dev /*[N,F] or NULL*/,
                        int N, int S, int F, float* out_dev /*[N,D]*/, void* stream);
int dg_emb_destroy(dg_emb* h);
/* Variant B -- pyannote/wespeaker-voxceleb-resnet34-LM (reference README.md:172-173, loaded through src/diart/models.py:50,59):
 * dg_emb_create recognises the checkpoint by its key names (resnet.conv1.weight, resnet.layer1.0.conv1.weight, ...,
 * resnet.seg_1.weight) and every dg_emb_* entry point then runs kaldi fbank -> ResNet34 -> TSTP -> Linear(5120, 256);
 * chunk lengths must be multiples of 160 samples.  Test hook: the trunk up to the stem (stop_after = -1) or BasicBlock
 * stop_after (0..15), returned as float32 [U][W][H][C] (time, mel, channel) on the host; -2 = log-mel features [U][T][80].
 * dims receives {U, W, H, C}.  Synchronises the device. */
int dg_emb_debug_trunk(dg_emb* h, const float* wav_dev, int U, int S, int stop_after, float* out_host, int64_t cap, int* dims);

/* ---- element-wise blocks ---- */
/* OverlappedSpeechPenalty (reference src/diart/blocks/embedding.py:98-107, functional.py:6-13) */
int dg_osp(const float* seg_dev /*[B,F,K]*/, int B, int F, int K, float gamma, float beta,
           int normalize, float* out_dev, void* stream);
/* EmbeddingNormalization (reference src/diart/functional.py:16-27): out = norm * e / ||e||_2 */
int dg_normalize_embeddings(const float* emb_dev /*[rows,D]*/, int rows, int D, float norm,
                            float* out_dev, void* stream);

/* ---- OnlineSpeakerClustering (reference src/diart/blocks/clustering.py:31-218 and the
 *      SpeakerMap logic of src/diart/mapping.py:179-360).  State (centroids, float64) lives on
 *      the device. ---- */
int dg_cluster_create(int max_speakers, int dim, double tau_active, double rho_update,
                      double delta_new, int device, dg_cluster** out);
/* Distance of the embeddings to the centroids: the `metric` argument of the reference class (clustering.py:31-46), handed to
 * scipy's cdist at mapping.py:175.  0 cosine (default), 1 euclidean, 2 sqeuclidean, 3 cityblock, 4 chebyshev; float64. */
int dg_cluster_set_metric(dg_cluster* h, int metric);
/* Processes the B chunks in order (the reference's sequential loop, diarization.py:193-203).
 * map_dev  int32 [B,K]: global speaker of each local speaker, -1 if unmapped.
 * permuted_dev float32 [B,F,M] or NULL: SpeakerMap.apply output (values are float32-exact). */
int dg_cluster_step(dg_cluster* h, const float* seg_dev /*[B,F,K]*/, const float* emb_dev /*[B,K,D]*/,
                    int B, int F, int K, int32_t* map_dev, float* permuted_dev, void* stream);
int dg_cluster_reset(dg_cluster* h);
/* synchronous host copies of the state: centers [M,D] float64, active [M] int32 (0/1);
 * *initialized = 0 until the first chunk was seen (reference `centers is None`). */
int dg_cluster_get_state(dg_cluster* h, double* centers_host, int32_t* active_host, int* initialized);
int dg_cluster_set_state(dg_cluster* h, const double* centers_host, const int32_t* active_host, int initialized);
int dg_cluster_destroy(dg_cluster* h);

/* ---- fused pipeline step: SpeakerDiarization.__call__ lines 177-203
 *      (reference src/diart/blocks/diarization.py): segmentation -> OSP -> embedding ->
 *      normalisation -> clustering, no host round trips.  The pipeline borrows the three handles. ---- */
int dg_pipeline_create(dg_seg* seg, dg_emb* emb, dg_cluster* clu, float gamma, float beta,
                       int normalize_weights, dg_pipeline** out);
int dg_pipeline_step(dg_pipeline* h, const float* wav_dev /*[B,S]*/, int B, int S,
                     float* seg_dev /*[B,F,K]*/, float* emb_dev /*[B,K,D]*/,
                     int32_t* map_dev /*[B,K]*/, float* permuted_dev /*[B,F,M] or NULL*/, void* stream);
/* Hint: consecutive windows of a batch are `hop_samples` apart in ONE stream (reference config.step x sample_rate;
 * windows as `rearrange_audio_stream` emits them, src/diart/operators.py:44-100).  The pipeline then verifies the overlap on the device for every batch (bit comparison) and, when it holds, runs the
 * sinc layer once over the unique samples instead of once per window (DG_STREAM_SINC=0 disables this).  0 = no hint
 * (default).  Results never depend on the hint being right. */
int dg_pipeline_set_hop(dg_pipeline* h, int hop_samples);
/* Same with HOST buffers: H2D of the waveforms and D2H of the results inside the call
 * (pinned staging owned by the handle); synchronous. */
int dg_pipeline_step_host(dg_pipeline* h, const float* wav_host, int B, int S, float* seg_host,
                          float* emb_host, int32_t* map_host, float* permuted_host /*nullable*/);
/* Pipelined variants for throughput: submit enqueues a step and returns; the sequential clustering of step i and
 * the host<->device copies overlap the networks of step i+1 (chunk order per stream is kept: all clustering runs on
 * one internal stream).  At most THREE steps may be outstanding -- two compute concurrently, the third lets the
 * waveforms of step i+2 be uploaded meanwhile -- and collect returns them oldest first.  dg_pipeline_collect makes
 * `stream` wait for the step and hands out device pointers that stay valid until the third next submit;
 * dg_pipeline_collect_host copies to host buffers and blocks. */
int dg_pipeline_submit(dg_pipeline* h, const float* wav_dev /*[B,S], must stay valid until collected*/, int B, int S,
                       void* stream);
int dg_pipeline_collect(dg_pipeline* h, const float** seg_dev, const float** emb_dev, const int32_t** map_dev,
                        void* stream);
/* as dg_pipeline_collect, but copies the step's results into caller-owned device buffers on `stream` */
int dg_pipeline_collect_copy(dg_pipeline* h, float* seg_dev, float* emb_dev, int32_t* map_dev, void* stream);
int dg_pipeline_submit_host(dg_pipeline* h, const float* wav_host /*pinned memory recommended*/, int B, int S);
int dg_pipeline_collect_host(dg_pipeline* h, float* seg_host, float* emb_host, int32_t* map_host);
int dg_pipeline_destroy(dg_pipeline* h);

/* ---- post-path of SpeakerDiarization.__call__ on the device (reference src/diart/blocks/diarization.py:205-232):
 *      SpeakerMap.apply (mapping.py:341-360) -> DelayedAggregation(step, latency, "hamming", "loose")
 *      (blocks/aggregation.py:73-92,120-218, incl. the first-buffer prepend rule :188-212) -> Binarize(tau)
 *      (blocks/utils.py:11-59), run-length encoded.  The handle keeps the scores / maps of the last num_windows - 1
 *      chunks (the reference's pred_buffer) on the device; num_windows = round(latency / step).
 *      hamming_host = np.hamming(frames) in float64.  Arithmetic is float64 in numpy's order without fused
 *      multiply-add: the thresholded result is bit-identical to the reference's.
 *
 *      plan_host int32 [B][4 + num_windows], one row per chunk, computed by the host with pyannote.core's
 *      SlidingWindow.crop index arithmetic (diart_b200/blocks/post.py):
 *        [0] nb   buffers aggregated for this chunk (1 .. num_windows)
 *        [1] nf   frames of the aggregated region
 *        [2] first_nf  > 0 only for the first buffer of a stream: output = first_nf frames, the last nf aggregated
 *        [3] first_lo  first frame of that prepended crop (may be negative: edge-padded)
 *        [4 + j]  first frame of buffer j's crop (oldest buffer first; may leave [0, frames): edge-padded)
 *      header_host int32 [B][4] = {offset into turns, number of turns, output frames, 0};
 *      turns_host  uint32 [turn_cap_host], packed speaker << 20 | on << 10 | off (frame indices; the turn covers the
 *      frame MIDDLES on .. off as in Binarize), each chunk's turns contiguous, by speaker then time.
 *      dg_post_step synchronises `stream`. ---- */
int dg_post_create(int frames, int local_speakers, int max_speakers, int num_windows, const double* hamming_host,
                   double tau, int device, dg_post** out);
int dg_post_step(dg_post* h, const float* seg_dev /*[B,F,K]*/, const int32_t* map_dev /*[B,K]*/, int B,
                 const int32_t* plan_host, int32_t* header_host, uint32_t* turns_host, int turn_cap_host, int* n_turns,
                 void* stream);
int dg_post_reset(dg_post* h);
int dg_post_destroy(dg_post* h);
/* The whole body of SpeakerDiarization.__call__ (reference diarization.py:172-232) in ONE call: rows_host[b] points to the
 * S float32 samples of window b (B separate host arrays, as rearrange_audio_stream emits them).  With a hop set
 * (dg_pipeline_set_hop) worker threads compare every window with its predecessor (memcmp of the S - hop shared samples);
 * windows that are consecutive hops of one stream are uploaded ONCE (S + (B-1) hop samples) and formed on the device,
 * anything else is gathered into pinned staging and uploaded as [B,S] while the gather is still running.  Then fused step
 * + post-path in up to three pipelined sub-batches; only the turn list (and, if asked for, scores and maps) returns to
 * the host.  Synchronous.  dg_pipeline_last_call_h2d_bytes: what the last call uploaded. */
int dg_pipeline_call_host(dg_pipeline* h, dg_post* post, const float* const* rows_host, int B, int S,
                          const int32_t* plan_host, int32_t* header_host, uint32_t* turns_host, int turn_cap_host,
                          int* n_turns, float* seg_host /*nullable*/, int32_t* map_host /*nullable*/);
int64_t dg_pipeline_last_call_h2d_bytes(const dg_pipeline* h);

/* ---- device-side audio stream: rearrange_audio_stream (reference src/diart/operators.py:44-100) with the ring buffer in
 *      HBM.  The host pushes every sample ONCE (step_samples new samples per chunk instead of chunk_samples: 8.2 MB
 *      instead of 82 MB per 256-chunk step at 5 s / 0.5 s); window i of the stream is samples
 *      [i * step_samples, i * step_samples + chunk_samples).  max_windows = largest batch that will be requested.
 *      dg_stream_push_host may be called with any block size (the reference's sources emit arbitrary blocks); it
 *      fails with DG_EINVAL when the ring is full, i.e. windows must be consumed first. ---- */
int dg_stream_create(int chunk_samples, int step_samples, int max_windows, int device, dg_stream** out);
int dg_stream_push_host(dg_stream* h, const float* samples_host, int n);
/* complete windows pushed but not yet consumed */
int dg_stream_available(const dg_stream* h);
/* the next B windows as a dense [B, chunk_samples] device batch on `stream`; advances the stream by B steps */
int dg_stream_windows(dg_stream* h, int B, float* wav_dev, void* stream);
int dg_stream_reset(dg_stream* h);
int dg_stream_destroy(dg_stream* h);
/* dg_pipeline_submit_host / dg_pipeline_call_host whose batch is the next B windows of `stream` (no window upload; the
 * sinc layer takes its stream form directly: the windows overlap by construction).  Collect with dg_pipeline_collect*. */
int dg_pipeline_submit_stream(dg_pipeline* h, dg_stream* stream, int B);
int dg_pipeline_call_stream(dg_pipeline* h, dg_post* post, dg_stream* stream, int B, const int32_t* plan_host,
                            int32_t* header_host, uint32_t* turns_host, int turn_cap_host, int* n_turns,
                            float* seg_host /*nullable*/, int32_t* map_host /*nullable*/);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t dg_launch_count(void);
/* per-kernel CUDA-event timing on the launching stream (bench.py's roofline leg).  While enabled,
 * every kernel launch is bracketed by two events; dg_profile_report() synchronises the device and
 * writes {"kernel": {"count": n, "ms": total}, ...} into buf. */
int dg_profile_enable(int enable);
/* test hook: the same random shifted-window GEMM through the float32 SIMT kernel and the tcgen05 (bf16x3)
 * kernel; epi 0 = bias -> f32, 1 = bias+leaky+bn -> bf16 hi/lo planes, 2 = bias+leaky+bn -> f32. */
int dg_selftest_gemm_tc(int M, int Cin, int KW, int dil, int N, int epi, float* max_abs_diff, float* out_rms);
/* test hook (host only, no GPU): the weight-side split of float32 values into the two 16-bit operand planes
 * (hi = rn16(x), lo = rn16(x - hi)); f16 = 1 -> IEEE fp16 (saturating), 0 -> bf16.  What the device does to
 * activations with cvt.rn(.satfinite).f16/bf16.f32. */
int dg_selftest_split_host(const float* x, long long n, int f16, unsigned short* hi, unsigned short* lo);
int dg_profile_report(char* buf, int cap);

/* ---- shared-identity mode (extension beyond the reference; SURVEY.md 8(e), BASELINE config 5): G ranks diarize
 *      independent streams against one table of global speakers.  Per pipeline step and rank:
 *        dg_cluster_export_delta -> record [M*D payload | M kinds | 2 reserved] float64 (what changed since the
 *        last merge); the host all-gathers the records (one NCCL all-gather, ~82 KB per rank);
 *        dg_cluster_merge applies all records in rank order with one deterministic rule (cluster.cu) so that all
 *        ranks hold bit-identical tables, and rewrites this rank's speaker maps of the step (maps_dev, n_maps
 *        int32 values) where a centre it created was merged into / moved to another index. ---- */
int dg_cluster_record_len(const dg_cluster* h);
int dg_cluster_export_delta(dg_cluster* h, double* record_dev, void* stream);
int dg_cluster_merge(dg_cluster* h, const double* records_dev /*[world, record_len]*/, int world, int rank,
                     int32_t* maps_dev /*nullable*/, int n_maps, void* stream);
/* The same exchange inside the pipelined flow (dg_pipeline_submit* / collect*): both calls are stream-ordered on the
 * pipeline's clustering stream, behind the clustering of every submitted step and ahead of the next one, so the protocol is
 * exactly the one-step-at-a-time protocol while the networks of the following steps keep running.
 *   dg_pipeline_identity_export: record_dev [record_len] float64 receives this rank's changes; `stream` waits for it
 *     (issue the all-gather on `stream`);
 *   dg_pipeline_identity_merge: the clustering stream waits for `stream`, merges, and relabels the maps of the steps
 *     clustered since the previous merge in their device slots (call the pair after every submit, before that step's collect). */
int dg_pipeline_identity_export(dg_pipeline* h, double* record_dev, void* stream);
int dg_pipeline_identity_merge(dg_pipeline* h, const double* records_dev, int world, int rank, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIART_B200_H */
