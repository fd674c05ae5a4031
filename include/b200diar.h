/* b200diar.h -- C ABI of the B200-native community-1 diarization hot path.
 *
 * Drop-in boundary for pyannote.audio's sliding-window inference path.  The reference is 100 % Python and has no
 * FFI for this path, so each entry point cites the *Python* interface it replaces (paths relative to
 * /root/reference/src/pyannote/audio).  All functions return 0 on success or a negative b200_status; the message
 * for the calling thread's last failure is available from b200_last_error().  No exceptions cross this boundary,
 * no torch types appear in it: device buffers are raw CUDA pointers owned by the caller, `stream` is a
 * cudaStream_t passed as void*, weights are host fp32 arrays in PyTorch state-dict layout.
 * One ctx per (process, device); a ctx is not thread-safe, use one per stream/thread.
 */
#ifndef B200DIAR_H_
#define B200DIAR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;

enum b200_status {
  B200_STATUS_OK = 0,
  B200_STATUS_INVALID = -1, /* bad argument / unsupported shape            -> ValueError  */
  B200_STATUS_CUDA = -2,    /* CUDA runtime / driver failure               -> RuntimeError */
  B200_STATUS_OOM = -3,     /* cudaErrorMemoryAllocation (inference.py:201-206 maps OOM to MemoryError) */
  B200_STATUS_STATE = -4    /* weights not loaded, ctx misuse                -> RuntimeError */
};

/* fixed geometry of the path */
#define B200_CHUNK_SAMPLES 160000
#define B200_FRAMES_PER_CHUNK 589
#define B200_LOCAL_SPEAKERS 3
#define B200_POWERSET_CLASSES 7
#define B200_EMBED_DIM 256
#define B200_FBANK_FRAMES 998
#define B200_MEL_BINS 80

const char* b200_last_error(void);
int b200_version(void);

/* Model.to(device) / Inference.to(device)  (core/inference.py:169-180) */
int b200_ctx_create(b200_ctx** ctx, int device);
int b200_ctx_destroy(b200_ctx* ctx);
/* Tuning / A-B options (no reference counterpart; results do not depend on the sub-batch sizes):
 *   "seg_max_batch" (4736) / "emb_max_batch" (296): chunks per sub-batch = workspace size;
 *   "conv_impl" 8 = per-layer choice of the tcgen05 conv kernels (default), 0 = CUDA-core reference conv, 1 = per-tap,
 *   2 = tcgen05 for stride-1 only, 3..6 = strip-streaming variants; "conv_fuse", "conv_fold", "conv_scfold", "conv_ghost";
 *   "seg_gemm_impl" / "seg_rec_impl" 1 = tensor cores, 0 = fp32 CUDA-core twins; "seg_conv_impl" 0 twins, 1 tensor
 *   cores, 2 sinc layer only, 3 Conv1d layers only; "fbank_share" 1 = overlapping chunks share their fbank frames;
 *   "profile" 1 = CUDA-event timers around the trunk / the segmentation (b200_ctx_timer).  Unknown keys and values out
 *   of range return B200_ERR_INVALID. */
int b200_ctx_set_option(b200_ctx* ctx, const char* key, int64_t value);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches claim) */
int64_t b200_ctx_launch_count(const b200_ctx* ctx);
/* with option "profile" = 1 the library brackets the ResNet trunk ("trunk": conv kernels) and the segmentation
 * network ("seg") with CUDA events on the caller's stream; this returns and resets the accumulated device time
 * and the number of units (segments / chunks) processed.  Synchronises the device. */
int b200_ctx_timer(b200_ctx* ctx, const char* name, double* total_ms, int64_t* units);

/* ---- weights: Model.from_pretrained state_dict (core/model.py:497-655) ------------------------------------ */

/* PyanNet (models/segmentation/PyanNet.py:92-161, models/blocks/sincnet.py:41-79). Index of LSTM arrays =
 * layer * 2 + direction (0 = forward, 1 = "_reverse"); PyTorch layouts ([4H][I], [4H][H], [4H]), gate order i,f,g,o. */
typedef struct b200_seg_weights {
  float wav_norm_weight, wav_norm_bias;       /* sincnet.wav_norm1d.{weight,bias}                       */
  const float* sinc_filters;                  /* [80][251] realised ParamSincFB bank (cos 0..39, sin 40..79) */
  const float* norm_weight[3];                /* sincnet.norm1d.{0,1,2}.weight  (80, 60, 60)            */
  const float* norm_bias[3];
  const float* conv_weight[2];                /* sincnet.conv1d.{1,2}.weight  [60][80][5], [60][60][5]  */
  const float* conv_bias[2];
  int32_t lstm_layers;                        /* <= 4 */
  const float* lstm_w_ih[8];
  const float* lstm_w_hh[8];
  const float* lstm_b_ih[8];
  const float* lstm_b_hh[8];
  const float* linear_weight[2];              /* linear.{0,1}.weight [128][256], [128][128]             */
  const float* linear_bias[2];
  const float* classifier_weight;             /* [7][128] */
  const float* classifier_bias;               /* [7]      */
} b200_seg_weights;
/* Model.load_state_dict / Model.from_pretrained (core/model.py:497-655) for PyanNet: host fp32 arrays in PyTorch layouts;
 * fp16 (hi, lo) splits, LSTM shared-memory images and the sinc bank's tensor-core layout are made here, once. */
int b200_seg_load(b200_ctx* ctx, const b200_seg_weights* w);

/* WeSpeakerResNet34 (models/embedding/wespeaker/resnet.py:84-145, 214-252).  conv weight [Cout][Cin][k][k] fp32,
 * eval-mode BatchNorm2d given by (weight, bias, running_mean, running_var), eps 1e-5; folded by the library. */
typedef struct b200_conv_bn {
  const float* conv_weight;                   /* NULL => layer absent (identity shortcut)               */
  const float* bn_weight;
  const float* bn_bias;
  const float* bn_mean;
  const float* bn_var;
} b200_conv_bn;
typedef struct b200_emb_weights {
  b200_conv_bn stem;                          /* resnet.conv1 / resnet.bn1                               */
  b200_conv_bn block_conv1[16];               /* resnet.layer{1..4}.{i}.conv1/bn1, blocks in order 3+4+6+3 */
  b200_conv_bn block_conv2[16];
  b200_conv_bn block_shortcut[16];            /* resnet.layer{2..4}.0.shortcut.{0,1}                      */
  const float* seg1_weight;                   /* resnet.seg_1.weight [256][5120]                          */
  const float* seg1_bias;                     /* [256]                                                    */
} b200_emb_weights;
/* the same for WeSpeakerResNet34 (models/embedding/wespeaker/__init__.py:324-372): eval-mode BatchNorm is folded into
 * the conv weights / biases, conv weights go to fp16 [tap][c_out][c_in] plus the per-kernel re-layouts. */
int b200_emb_load(b200_ctx* ctx, const b200_emb_weights* w);

/* ---- audio ingest: Audio.__call__ / Audio.downmix_and_resample (core/io.py:223-265, 306-351) -----------------
 * pcm is a DEVICE buffer holding the raw decoded audio: B200_PCM_S16_INTERLEAVED = int16 [frame][channel] (what a
 * PCM WAV holds; half the PCIe bytes of float32) or B200_PCM_F32_PLANAR = float32 [channel][frame] (the reference's
 * in-memory {"waveform": (channel, time)} files).  channel >= 0 selects that channel (io.py:232-233), channel < 0
 * downmixes by the mean over channels (mono="downmix", io.py:241-242); when sr_in != sr_out the signal is resampled
 * with torchaudio.functional.resample's default polyphase windowed-sinc filter (io.py:246-250).  Writes
 * b200_audio_num_frames(frames_in, sr_in, sr_out) float32 samples to out (capacity checked). */
#define B200_PCM_S16_INTERLEAVED 0
#define B200_PCM_F32_PLANAR 1
int64_t b200_audio_num_frames(int64_t frames_in, int32_t sr_in, int32_t sr_out);
int b200_audio_ingest(b200_ctx* ctx, const void* pcm, int32_t format, int32_t channels, int64_t frames_in,
                      int32_t sr_in, int32_t sr_out, int32_t channel, float* out, int64_t out_capacity, void* stream);

/* ---- segmentation: Inference.infer / Inference.slide hot loop (core/inference.py:182-215, 295-313) --------
 * `wav` is a device fp32 buffer; chunk i covers wav[chunk_off[i] .. chunk_off[i]+160000), of which only the first
 * chunk_valid[i] samples are real (the rest is the zero padding of the last chunk, inference.py:270-278).
 * chunk_off / chunk_valid are HOST arrays.  Output: powerset class id per frame, classes[num_chunks][589]
 * (argmax of the LogSoftmax output, utils/powerset.py:135-140); optional log-probabilities [num_chunks][589][7]. */
int b200_seg_forward(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                     int32_t num_chunks, uint8_t* classes, float* logp, void* stream);
/* SincNet.forward alone (models/blocks/sincnet.py:163-184): out[num_chunks][589][60] fp32 (frame-major). */
int b200_sincnet_forward(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                         int32_t num_chunks, float* out, void* stream);
/* Powerset.to_multilabel, hard (utils/powerset.py:115-140): classes[n] -> multilabel[n][3] in {0,1} (u8). */
int b200_powerset_to_multilabel(b200_ctx* ctx, const uint8_t* classes, int64_t n, uint8_t* multilabel, void* stream);

/* ---- embeddings: SpeakerDiarization.get_embeddings hot loop (pipelines/speaker_diarization.py:399-459) over
 * PyannoteAudioPretrainedSpeakerEmbedding.__call__ (pipelines/speaker_verification.py:704-716) and
 * WeSpeakerResNet34.forward (models/embedding/wespeaker/__init__.py:324-343).  One trunk pass per chunk, the three
 * local speakers share it (forward_frames + forward_embedding, :288-322); masks[num_chunks][3][589] u8 are the
 * StatsPool weights; emb[num_chunks][3][256] fp32. */
int b200_emb_forward(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                     int32_t num_chunks, const uint8_t* masks, float* emb, void* stream);
/* The same with a fused all-gather for the multi-GPU chunk pool: emb_peers[n_peers] (HOST array
 * of DEVICE pointers, n_peers <= 7) are this rank's slot inside the OTHER GPUs' gather buffers (peer memory mapped
 * over NVLink, e.g. CUDA IPC / torch symmetric memory); the epilogue of the final Linear GEMM stores every output
 * tile to `emb` and to all peers (P2P stores), so the exchange overlaps the GEMM and no collective call follows.
 * The caller synchronises the ranks afterwards (any barrier with system-scope release/acquire). */
int b200_emb_forward_push(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                          int32_t num_chunks, const uint8_t* masks, float* emb, float* const* emb_peers,
                          int32_t n_peers, void* stream);
/* P2P push of a byte range (16-byte aligned) to the same offsets of n_dsts <= 7 peer buffers: the powerset classes
 * of this rank's chunks, next to the embeddings pushed by b200_emb_forward_push. */
int b200_push(b200_ctx* ctx, const void* src, int64_t bytes, void* const* dsts, int32_t n_dsts, void* stream);
/* compute_fbank (wespeaker/__init__.py:113-139): fbank[num_chunks][998][80], global-mean centred. */
int b200_emb_fbank(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                   int32_t num_chunks, float* fbank, void* stream);
/* Host-only (no device work): the shared-frame fbank layout b200_emb_forward uses for a chunk list processed in
 * sub-batches of `sub_batch` chunks.  The reference computes 998 fbank frames per chunk (wespeaker/__init__.py:113-139)
 * although consecutive chunks of Inference.slide (core/inference.py:235-257: step = 0.1 * duration = 100 frame hops)
 * share 898 of them; overlapping, hop-aligned, full chunks are grouped into runs whose frames are computed once.
 * frame0[num_chunks]: first fbank row of each chunk inside its sub-batch; rows_per_sub_batch[ceil(n / sub_batch)]:
 * fbank rows computed per sub-batch (998 * chunks without sharing).  Returns the number of runs, < 0 on bad arguments. */
int64_t b200_emb_fbank_plan(const int64_t* chunk_off, const int32_t* chunk_valid, int32_t num_chunks, int32_t sub_batch,
                            int32_t share, int32_t* frame0, int32_t* rows_per_sub_batch);
/* ResNet.forward_frames on a given fbank (resnet.py:399-419): frames[num_chunks][256][10][125] fp32 (NCHW). */
int b200_emb_trunk(b200_ctx* ctx, const float* fbank, int32_t num_chunks, float* frames, void* stream);
/* StatsPool.forward (models/blocks/pooling.py:76-130): seq[B][F][T], weights[B][S][Tw] or NULL -> out[B][S][2F]. */
int b200_stats_pool(b200_ctx* ctx, const float* seq, const float* weights, float* out, int32_t B, int32_t F, int32_t T,
                    int32_t S, int32_t Tw, void* stream);

/* ---- overlap-add / reconstruction (core/inference.py:498-620, pipelines/utils/diarization.py:150-268,
 * pipelines/speaker_diarization.py:480-528).  seg[num_chunks][589][3] u8 in {0,1}; start_frame[num_chunks] DEVICE
 * int32 array with the global frame index of each chunk's first frame (non-decreasing; computed on the host as
 * inference.py:596 does); num_frames = size of the global grid. */
int b200_speaker_count(b200_ctx* ctx, const uint8_t* seg, const int32_t* start_frame, int32_t num_chunks,
                       int32_t num_frames, uint8_t* count, void* stream);
/* SpeakerDiarization.reconstruct + to_diarization (pipelines/speaker_diarization.py:480-528,
 * pipelines/utils/diarization.py:221-268): per frame, the `count` most active clusters (ties: lower cluster index).
 * hard_clusters[num_chunks][3] int8 DEVICE (-2 = inactive/unassigned, values >= num_clusters_out are ignored);
 * count[num_frames] u8 device (already capped); out: discrete[num_frames][num_clusters_out] u8 with
 * num_clusters_out >= max(K, max(count)), at most 127 (hard clusters are int8 like the reference's
 * constrained_argmax; up to 32 clusters the per-frame counters stay in registers). */
int b200_reconstruct(b200_ctx* ctx, const uint8_t* seg, const int8_t* hard_clusters, const int32_t* start_frame,
                     int32_t num_chunks, int32_t num_frames, const uint8_t* count, int32_t num_clusters_out,
                     uint8_t* discrete, void* stream);

/* Inference.aggregate (core/inference.py:498-620), the generic float overlap-add behind the aggregated
 * (skip_aggregation=False) Inference output and the VAD / OSD pipelines: scores[num_chunks][589][K] fp32 (NaN =
 * missing), hamming / warm_up: DEVICE fp64[589] windows or NULL (= ones) -> out[num_frames][K] fp32, bit-identical to
 * numpy's mixed float32/float64 arithmetic (see post.cu). */
int b200_aggregate(b200_ctx* ctx, const float* scores, const int32_t* start_frame, int32_t num_chunks,
                   int32_t num_frames, int32_t num_classes, const double* hamming, const double* warm_up,
                   int32_t skip_average, float missing, float epsilon, float* out, void* stream);
/* VoiceActivityDetection's pre-aggregation step (pipelines/voice_activity_detection.py:111-114: max over the
 * speakers of the multilabel output) straight from the powerset classes: speech[n] fp32 in {0,1}. */
int b200_powerset_speech(b200_ctx* ctx, const uint8_t* classes, int64_t n, float* speech, void* stream);

/* Onsets / offsets of a discrete diarization discrete[num_frames][num_clusters] u8 (to_annotation ->
 * Binarize(onset=offset=0.5), pipelines/utils/diarization.py:188-218, utils/signal.py:254-318) as unordered events
 * k * (num_frames + 1) + f.  buf (DEVICE int32[2 + 2 * cap]) = [n_on, n_off, on[cap], off[cap]]; counts may exceed
 * cap (then only cap events were stored: call again with a larger buffer). */
int b200_frame_transitions(b200_ctx* ctx, const uint8_t* discrete, int32_t num_frames, int32_t num_clusters,
                           int32_t cap, int32_t* buf, void* stream);

/* ---- clustering (pipelines/clustering.py:77-140, 572-669; utils/vbx.py; scipy linkage/fcluster) --------------- */
/* filter_embeddings: clean-frame counts per (chunk, speaker): out[num_chunks][3] int32, plus active[num_chunks][3] u8
 * = any frame active (inactive speakers, speaker_diarization.py:681). */
int b200_clean_frames(b200_ctx* ctx, const uint8_t* seg, int32_t num_chunks, int32_t* clean, uint8_t* active,
                      void* stream);
/* linkage(X, "centroid", "euclidean") (scipy, called at clustering.py:600-602 and :374-376): x[n][dim] fp64 device;
 * normalize: 0 = rows as given, 1 = L2-normalised in fp64, 2 = rows hold float32 values and are normalised exactly as
 * numpy does on float32 embeddings (x / np.linalg.norm(x, axis=1, keepdims=True): float32 pairwise sum, float32
 * sqrt and division; clustering.py:597-599) before widening; Z[n-1][4] fp64 device in scipy's format.
 * Limits: n <= 32768 observations per problem (a dense n x n fp64 distance matrix lives in the ctx workspace: 8 n^2
 * bytes, 8.6 GB at the limit); longer recordings must be clustered in windows by the caller. */
int b200_linkage_centroid(b200_ctx* ctx, const double* x, int32_t n, int32_t dim, int32_t normalize, double* Z,
                          void* stream);
/* the same (clustering.py:594-603) for num_problems independent problems in ONE launch (one CTA each): rows of problem f are
 * x[row_offsets[f] .. row_offsets[f+1]) (row_offsets: HOST int32[num_problems+1]); Z rows are concatenated, problem
 * f contributing max(n_f - 1, 0) rows. */
int b200_linkage_centroid_batched(b200_ctx* ctx, const double* x, const int32_t* row_offsets, int32_t num_problems,
                                  int32_t dim, int32_t normalize, double* Z, void* stream);
/* fcluster(Z, t, criterion="distance") (clustering.py:604, 385): HOST arrays, labels[n] 1-based like scipy. */
int b200_fcluster_distance(const double* Z, int32_t n, double t, int32_t* labels);
/* PLDA.__call__ (core/plda.py:50-63; xvec_tf / plda_tf of utils/vbx.py:211-217): x[n][Din] fp64 device ->
 * fea[n][L]; mean1[Din], mean2[Dout], lda[Din][Dout], mu[Dout], trT[Dout][L] (= plda_tr^T[:, :L]) fp64 device. */
int b200_plda_transform(b200_ctx* ctx, const double* x, int32_t n, int32_t Din, int32_t Dout, int32_t L,
                        const double* mean1, const double* mean2, const double* lda, const double* mu,
                        const double* trT, double* fea, void* stream);
/* VBx centroids (clustering.py:620-621): W = q[:, kept]; centroids[K][dim] = W^T train / sum(W): q[n][S], kept[K]
 * (DEVICE int32 column indices), train[n][dim], all fp64 device. */
int b200_weighted_centroids(b200_ctx* ctx, const double* q, int32_t n, int32_t S, const int32_t* kept, int32_t K,
                            const double* train, int32_t dim, double* centroids, void* stream);
/* cdist(a, b, "cosine") (clustering.py:645-655): a[m][dim], b[k][dim] fp64 device -> d[m][k] fp64 device. */
int b200_cdist_cosine(b200_ctx* ctx, const double* a, int32_t m, const double* b, int32_t k, int32_t dim, double* d,
                      void* stream);
/* VBx iterations (utils/vbx.py:98-136 via cluster_vbx :140-155): fea[n][D], phi[D], gamma[n][S] (in: initial
 * responsibilities, out: final), pi[S] (out), all fp64 device; *iters (host) = iterations run. */
int b200_vbx(b200_ctx* ctx, const double* fea, const double* phi, int32_t n, int32_t D, int32_t S, double Fa,
             double Fb, int32_t max_iters, double epsilon, double* gamma, double* pi, int32_t* iters, void* stream);
/* batched: problem f has n[f] frames (consecutive rows of fea) and S[f] speakers; gamma / pi are the per-problem
 * arrays concatenated; n, S, iters (nullable) are HOST int32[num_problems].  One 8-CTA thread-block cluster per
 * problem runs all iterations (utils/vbx.py:98-136), convergence is tested on the device. */
int b200_vbx_batched(b200_ctx* ctx, const double* fea, const double* phi, const int32_t* n, const int32_t* S,
                     int32_t num_problems, int32_t D, double Fa, double Fb, int32_t max_iters, double epsilon,
                     double* gamma, double* pi, int32_t* iters, void* stream);
/* constrained_argmax / argmax (clustering.py:127-140, 658-665): soft[num_chunks][3][K] fp64 device ->
 * hard[num_chunks][3] int8 device (-2 = unassigned). */
int b200_assign(b200_ctx* ctx, const double* soft, int32_t num_chunks, int32_t num_clusters, int32_t constrained,
                int8_t* hard, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DIAR_H_ */
