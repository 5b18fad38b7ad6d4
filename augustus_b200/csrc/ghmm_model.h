/*
 * ghmm_model.h — host side: AUGB2PAR blob -> quantised model tables (DevModel).
 *
 * The blob carries what NAMGene::NAMGene() and StateModel::readAllParameters() leave in the
 * reference's statics (namgene.cc:25-142, 1318-1392; exonmodel.cc:604-792; intronmodel.cc:295-415;
 * igenicmodel.cc:150-225).  Everything is converted once to Q40 fixed point.
 */
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/augb200.h"
#include "../../include/augb200_params.h"
#include "ghmm_defs.h"

namespace augb {

inline sc_t quantize(double x) {
    if (!(x > -1e300)) return SC_NEG;
    return (sc_t)llround(ldexp(x, FRAC_BITS));
}

struct HostModel {
    DevModel dm;                       /* pointers refer to `tab` (host) */
    std::vector<sc_t> tab;             /* all tables, contiguous */
    std::vector<size_t> off;           /* offsets (in sc_t units) of the pointer members, in fixed order */
    std::string err;

    struct BlobView {
        const char* p; size_t n;
        const augb200_blob_entry* find(const char* name) const {
            const augb200_blob_header* h = (const augb200_blob_header*)p;
            const augb200_blob_entry* e = (const augb200_blob_entry*)(p + sizeof *h);
            for (uint32_t i = 0; i < h->n_entries; i++) if (!strncmp(e[i].name, name, sizeof e[i].name) && strlen(name) < sizeof e[i].name) return &e[i];
            return nullptr;
        }
    };

    int build(const void* blob, size_t nbytes);
    /* copy of dm with table pointers rebased onto `base` (e.g. a device allocation holding tab) */
    DevModel rebased(const sc_t* base) const;
};

}  // namespace augb
