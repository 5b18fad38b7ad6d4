/*
 * augb200.h — C ABI of the B200-native GHMM decoder (libaugb200.so).
 *
 * This is the drop-in boundary for AUGUSTUS' DP engine.  Each entry point names the reference
 * interface (AUGUSTUS 3.5.0, paths relative to the reference tree) it replaces; INTEGRATION.md shows
 * the C++ shim a maintainer adds to NAMGene to call it.
 *
 *   augb200_model_create   <- NAMGene::NAMGene() + StateModel::readAllParameters()
 *                             (src/namgene.cc:25-142, src/augustus.cc:175-176): the host exports its
 *                             tables once as an AUGB2PAR blob (include/augb200_params.h).
 *   augb200_decode_batch   <- NAMGene::viterbiAndForward (src/namgene.cc:168-365) followed by
 *                             NAMGene::getViterbiPath (src/namgene.cc:432-510), for many windows at
 *                             once; the StateModel::viterbiForwardAndSampling overrides
 *                             (exonmodel.cc:899, intronmodel.cc:509, igenicmodel.cc:231) run inside
 *                             as CUDA kernels.
 *   augb200_decode         <- the same for one window (what findGenes calls, namgene.cc:776,790).
 *   augb200_strerror       <- the ProjectError texts thrown at namgene.cc:455-457 / :493-496.
 *
 * There is no CPU fallback: every entry point that needs the device fails with
 * AUGB200_ERR_CUDA / AUGB200_ERR_NO_DEVICE when no sm_100 GPU is usable.
 *
 * Plain C types only; all pointers are HOST pointers owned by the caller unless stated otherwise.
 */
#ifndef AUGB200_H
#define AUGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUGB200_VERSION 1

/* error codes (0 = success) */
enum {
    AUGB200_OK = 0,
    AUGB200_ERR_BAD_BLOB = 1,        /* malformed / incomplete parameter blob                         */
    AUGB200_ERR_UNSUPPORTED = 2,     /* model uses states / options this library does not decode      */
    AUGB200_ERR_NO_DEVICE = 3,       /* no CUDA device                                                */
    AUGB200_ERR_CUDA = 4,            /* CUDA runtime error (see augb200_last_cuda_error)              */
    AUGB200_ERR_BAD_ARG = 5,
    AUGB200_ERR_NO_PATH = 6,         /* "No feasible path found in HMM" (namgene.cc:455-457)          */
    AUGB200_ERR_STUCK = 7,           /* "Viterbi got stuck" (namgene.cc:493-496)                      */
    AUGB200_ERR_CAPACITY = 8,        /* internal event capacity exceeded for a window                 */
    AUGB200_ERR_TOO_LONG = 9         /* window longer than AUGB200_MAX_WINDOW                         */
};

/* longest window the Q40 fixed-point score range covers */
#define AUGB200_MAX_WINDOW 500000

/* truncation flags of a path state, as State::truncated (include/types.hh:69-70) */
#define AUGB200_TRUNC_LEFT 1
#define AUGB200_TRUNC_RIGHT 2

typedef struct augb200_model augb200_model;

/* One sequence window = one call of NAMGene::viterbiAndForward. */
typedef struct augb200_window {
    const char*    dna;       /* ASCII; anything but acgt/ACGT is an unknown base.  Case matters only for models exported with
                                  softmasking on: a lower-case base is soft-masked (SequenceFeatureCollection::prepare,
                                  extrinsicinfo.cc:1696-1724) and gets the nonexonpart bonus in every non-exon state */
    int32_t        length;    /* number of bases, 2 .. AUGB200_MAX_WINDOW                          */
    const int32_t* gc_class;  /* optional: ContentStairs::idx[] (motif.cc:543-614), one class per  */
                              /* base; NULL = computed on the device with the same rule            */
} augb200_window;

/*
 * A decoded state path: StatePath (include/gene.hh) in left-to-right order with runs of identical
 * non-exon states already merged as StatePath::condenseStatePath does (gene.cc:977-1000).
 * type[] holds StateType values (include/types.hh:492-512); begin/end are 0-based inclusive.
 * The arrays are owned by the library and stay valid until the next decode call on the same
 * model or augb200_model_destroy.
 */
typedef struct augb200_path {
    int32_t        n;         /* number of states                                                  */
    int32_t        status;    /* AUGB200_OK or AUGB200_ERR_NO_PATH / _STUCK / _CAPACITY            */
    const int32_t* begin;
    const int32_t* end;
    const uint8_t* type;
    const uint8_t* truncated;
    double         log_prob;  /* ln of pathemiProb (Viterbi score incl. termProbs)                 */
} augb200_path;

/* Create a model from an AUGB2PAR blob (copied; the caller may free it afterwards). `device` is the
 * CUDA device ordinal.  Several models may live on one device and be used in turn by the host thread (each call uploads its own
 * model constants): the reference swaps initProbs / termProbs for the synch-state vectors at the cut points of long sequences
 * (NAMGene::doViterbiPiecewise, namgene.cc:594-603) — a host serves that with one model per vector pair (host/augshim.cc).
 * Synchronous calls on models of the same device take turns (one set of model constants per device). */
int augb200_model_create(const void* blob, size_t nbytes, int device, augb200_model** out);
void augb200_model_destroy(augb200_model* m);

/* number of GHMM states / GC classes of the model */
int augb200_model_statecount(const augb200_model* m);
int augb200_model_num_gc_classes(const augb200_model* m);

/* Decode n windows.  out[i] describes window i.  Returns AUGB200_OK if the batch ran; per-window
 * DP errors are reported in out[i].status. */
int augb200_decode_batch(augb200_model* m, int32_t n, const augb200_window* windows, augb200_path* out);

/*
 * Viterbi path plus posterior samples: NAMGene::findGenes with --sample=nsample (namgene.cc:833-871).  The forward matrix
 * is filled beside the Viterbi matrix and nsample-1 state paths are drawn as NAMGene::getSampledPath does
 * (namgene.cc:367-426).  Every window draws from its own rand() stream started at glibc's default seed, i.e. the stream an
 * `augustus` process running that window alone consumes.  samples[i*(nsample-1) + k] is sample k of window i (condensed
 * like the Viterbi path; log_prob = ln of the product of the normalised option probabilities, StatePath::pathemiProb).
 */
int augb200_decode_batch_sampling(augb200_model* m, int32_t n, const augb200_window* windows, int32_t nsample,
                                  augb200_path* out, augb200_path* samples);

/*
 * rand() stream position for hosts that predict several sequences in ONE process (the reference shares one glibc rand() stream,
 * seed 1, across all getSampledPath calls of a run, vitmatrix.cc:300): the windows of the next augb200_decode_batch_sampling call
 * start `draws_consumed` values into that stream (default 0 = a fresh process per window, the protocol of BASELINE.md), and
 * augb200_last_rand_consumed returns how many values the sampled paths of window 0 of the last call took.  A host that decodes
 * one window per call in input order and adds up the consumed counts reproduces a single reference process exactly.
 */
int augb200_set_rand_position(augb200_model* m, uint64_t draws_consumed);
int64_t augb200_last_rand_consumed(const augb200_model* m);

/* Decode one window. */
int augb200_decode(augb200_model* m, const augb200_window* window, augb200_path* out);

/*
 * Several devices, one call: `models` are n_models models created from the same blob on different devices; window i
 * is decoded by models[i mod n_models] (block-cyclic, like the reference's own scale-out: one `augustus --predictionStart/End` job
 * per chunk of scripts/createAugustusJoblist.pl), one host thread per device, no exchange between devices.  out[] (and samples[],
 * when nsample >= 2 asks for sampled paths as in augb200_decode_batch_sampling; pass nsample = 0 and samples = NULL otherwise)
 * come back in input order; the arrays of window i stay owned by the model that decoded it until its next call.  This is the
 * call a C++ host uses to spread the windows of a chromosome over the GPUs of a node; ranks of a multi-process job use
 * augb200_decode_batch per rank and gather the path arrays themselves (bench.py: one NCCL gather).
 */
int augb200_decode_batch_multi(augb200_model* const* models, int32_t n_models, int32_t n, const augb200_window* windows,
                               int32_t nsample, augb200_path* out, augb200_path* samples);

/* Device-resident variant used for kernel-only timing: stage the windows once, run the kernels any
 * number of times (no host<->device traffic in between), fetch the paths at the end.  DNA and window
 * descriptors of the whole batch stay on the device; the per-window workspaces come from one arena that
 * equal-sized waves of windows use in turn, so the batch may be larger than the device memory left for
 * workspaces (10 000 x 50 kb = two waves on a 180 GB part). */
int augb200_stage_batch(augb200_model* m, int32_t n, const augb200_window* windows);
int augb200_run_staged(augb200_model* m);                 /* asynchronous on the model's stream    */
int augb200_fetch_staged(augb200_model* m, augb200_path* out);
/* the CUDA stream (cudaStream_t) the kernels are launched on, for event timing by the caller */
void* augb200_model_stream(augb200_model* m);
/* number of kernel launches issued by the last run / decode call */
int64_t augb200_last_launch_count(const augb200_model* m);
/* device milliseconds spent in the sweep kernel during the last run (CUDA events on the model's stream) */
double augb200_last_sweep_ms(const augb200_model* m);

/* The contiguous store that every augb200_path of the last decode / fetch call points into (all windows'
 * states back to back); returns the number of states in it. */
int64_t augb200_result_store(const augb200_model* m, const int32_t** begin, const int32_t** end,
                             const uint8_t** type, const uint8_t** truncated);

/* the same for the sampled paths of the last augb200_decode_batch_sampling call */
int64_t augb200_sample_store(const augb200_model* m, const int32_t** begin, const int32_t** end,
                             const uint8_t** type, const uint8_t** truncated);

/*
 * De-duplicated sampled paths (first step of SURVEY.md §8f next-1).  NAMGene::findGenes compares the transcripts of every sampled path
 * with the ones it has collected (Transcript::operator==, namgene.cc:875-904); the library finds repeated state paths on the device: a
 * sample whose states equal those of an earlier sample of the same window shares that sample's arrays (its augb200_path points to the
 * same begin/end/type/truncated, so callers that walk every path see no difference), is not copied to the host a second time, and
 * first[i*(nsample-1) + k] is the index of the first sample of window i with the same states (k itself for a new path).  A host that
 * post-processes only the first occurrences and weighs them by their multiplicity does 1 / (duplicate factor) of the work.
 * Returns the number of states of all sampled paths of the last augb200_decode_batch_sampling call with duplicates counted
 * (augb200_sample_store returns the number actually held).
 */
int64_t augb200_sample_first_occurrence(const augb200_model* m, const int32_t** first);

const char* augb200_strerror(int code);
const char* augb200_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif
