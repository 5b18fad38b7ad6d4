/*
 * augb200_params.h — the parameter blob ("AUGB2PAR") consumed by augb200_model_create().
 *
 * The blob is a flat, position-independent byte buffer: header, table of named arrays, payload.
 * It carries every table the GHMM decoder needs, exported by the host AFTER
 * StateModel::readAllParameters() (reference src/augustus.cc:176) so that load-time blending /
 * tying (igenic<-intron, utrmodel.cc:680-688) is already applied.  All probabilities are natural
 * logarithms in FP64 (Double::log(), reference include/lldouble.hh:146-148) with -inf for exact
 * zeros; integers are int32.  The library quantises the logs to Q23.40 fixed point itself.
 *
 * Array names are listed in DESIGN.md ("parameter blob") and in oracle/augdump.cc, which writes
 * the blob from an unmodified reference build.
 */
#ifndef AUGB200_PARAMS_H
#define AUGB200_PARAMS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUGB200_BLOB_MAGIC "AUGB2PAR"
#define AUGB200_BLOB_VERSION 1u
#define AUGB200_DT_F64 0u
#define AUGB200_DT_I32 1u

typedef struct augb200_blob_header {
    char     magic[8];      /* "AUGB2PAR" */
    uint32_t version;       /* AUGB200_BLOB_VERSION */
    uint32_t n_entries;
} augb200_blob_header;

typedef struct augb200_blob_entry {
    char     name[40];      /* NUL-terminated */
    uint32_t dtype;         /* AUGB200_DT_* */
    uint32_t ndim;          /* 1..4 */
    uint64_t dims[4];
    uint64_t offset;        /* from start of blob, 64-byte aligned */
    uint64_t nbytes;
} augb200_blob_entry;

#ifdef __cplusplus
}
#endif
#endif
