/*
 * aug_export.h — host-side exporter of the parameter blob (shared by the drop-in binding host/augshim.cc and the test tool oracle/augdump.cc).
 *
 * Exports the AUGB2PAR parameter blob (include/augb200_params.h) from the reference's static tables after
 * StateModel::readAllParameters() (reference src/augustus.cc:176).  This is the function body a maintainer
 * moves into the host.
 *
 * Access to the reference's private statics: the including translation unit is compiled with
 * private/protected re-defined AFTER all standard headers are included; class layout is unaffected by
 * access specifiers under the Itanium ABI.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <climits>
#include <cctype>
#include <ctime>
#include <exception>
#include <stdexcept>
#include <iterator>
#include <queue>
#include <deque>
#include <stack>
#include <bitset>
#include <memory>
#include <functional>
#include <unordered_map>
#include <unordered_set>
#include <typeinfo>
#include <numeric>
#include <locale>
#include <cassert>

#define private public
#define protected public
#include "types.hh"
#include "gene.hh"
#include "genbank.hh"
#include "namgene.hh"
#include "evaluation.hh"
#include "statemodel.hh"
#include "exonmodel.hh"
#include "intronmodel.hh"
#include "igenicmodel.hh"
#include "utrmodel.hh"
#include "ncmodel.hh"
#include "extrinsicinfo.hh"
#include "motif.hh"
#include "geneticcode.hh"
#include "properties.hh"
#include "pp_scoring.hh"
#undef private
#undef protected

#include "augb200_params.h"

/* ---------------------------------------------------------------- blob writer */
struct BlobWriter {
    struct Ent { augb200_blob_entry e; std::vector<char> data; };
    std::vector<Ent> ents;
    void add(const char* name, uint32_t dtype, std::vector<uint64_t> dims, const void* p, size_t nbytes) {
        Ent x; memset(&x.e, 0, sizeof x.e);
        if (strlen(name) >= sizeof x.e.name) throw ProjectError(std::string("augb200 export: array name too long: ") + name);
        strcpy(x.e.name, name);
        x.e.dtype = dtype; x.e.ndim = (uint32_t)dims.size();
        uint64_t n = 1;
        for (size_t i = 0; i < dims.size(); i++) { x.e.dims[i] = dims[i]; n *= dims[i]; }
        size_t es = dtype == AUGB200_DT_F64 ? 8 : 4;
        if (n * es != nbytes) throw ProjectError(std::string("augb200 export: unexpected table size for ") + name + " (this model configuration is not exportable)");
        x.e.nbytes = nbytes;
        x.data.assign((const char*)p, (const char*)p + nbytes);
        ents.push_back(x);
    }
    void f64(const char* name, std::vector<uint64_t> dims, const std::vector<double>& v) {
        add(name, AUGB200_DT_F64, dims, v.data(), v.size() * 8);
    }
    void i32(const char* name, std::vector<uint64_t> dims, const std::vector<int32_t>& v) {
        add(name, AUGB200_DT_I32, dims, v.data(), v.size() * 4);
    }
    void scalar_i(const char* name, int v) { std::vector<int32_t> x(1, v); i32(name, {1}, x); }
    void scalar_f(const char* name, double v) { std::vector<double> x(1, v); f64(name, {1}, x); }
    std::vector<char> bytes() {
        augb200_blob_header h; memcpy(h.magic, AUGB200_BLOB_MAGIC, 8);
        h.version = AUGB200_BLOB_VERSION; h.n_entries = (uint32_t)ents.size();
        uint64_t off = sizeof h + ents.size() * sizeof(augb200_blob_entry);
        for (auto& x : ents) { off = (off + 63) & ~63ull; x.e.offset = off; off += x.e.nbytes; }
        std::vector<char> out(off, 0);
        memcpy(out.data(), &h, sizeof h);
        for (size_t i = 0; i < ents.size(); i++) memcpy(out.data() + sizeof h + i * sizeof(augb200_blob_entry), &ents[i].e, sizeof(augb200_blob_entry));
        for (auto& x : ents) memcpy(out.data() + x.e.offset, x.data.data(), x.data.size());
        return out;
    }
    void write(const char* fn) {
        augb200_blob_header h; memcpy(h.magic, AUGB200_BLOB_MAGIC, 8);
        h.version = AUGB200_BLOB_VERSION; h.n_entries = (uint32_t)ents.size();
        uint64_t off = sizeof h + ents.size() * sizeof(augb200_blob_entry);
        for (auto& x : ents) { off = (off + 63) & ~63ull; x.e.offset = off; off += x.e.nbytes; }
        FILE* f = fopen(fn, "wb");
        if (!f) { perror(fn); exit(2); }
        fwrite(&h, sizeof h, 1, f);
        for (auto& x : ents) fwrite(&x.e, sizeof x.e, 1, f);
        uint64_t pos = sizeof h + ents.size() * sizeof(augb200_blob_entry);
        for (auto& x : ents) {
            while (pos < x.e.offset) { fputc(0, f); pos++; }
            fwrite(x.data.data(), 1, x.data.size(), f); pos += x.data.size();
        }
        fclose(f);
    }
};

static double lg(Double x) {               // Double::log(), -inf for exact zero
    if (!(x > Double(0.0))) return -std::numeric_limits<double>::infinity();
    return x.log();
}
static std::vector<double> lgv(const std::vector<Double>& v) {
    std::vector<double> r(v.size());
    for (size_t i = 0; i < v.size(); i++) r[i] = lg(v[i]);
    return r;
}

static void export_motif(BlobWriter& bw, const char* name, Motif* arr, int C) {
    int n = arr[0].n, k = arr[0].k;
    size_t w = n > 0 ? arr[0].windowProbs[0].size() : 0;        // a motif of width 0 (toxoplasma: no TATA / TTS motif) has no windowProbs
    std::vector<double> v; v.reserve((size_t)C * n * w);
    for (int c = 0; c < C; c++) {
        if (arr[c].n != n || arr[c].k != k) throw ProjectError(std::string("augb200 export: motif shape differs across GC classes: ") + name);
        for (int i = 0; i < n; i++) { auto t = lgv(arr[c].windowProbs[i]); v.insert(v.end(), t.begin(), t.end()); }
    }
    bw.f64(name, {(uint64_t)C, (uint64_t)n, (uint64_t)w}, v);
    std::string s(name); bw.scalar_i((s + "_n").c_str(), n); bw.scalar_i((s + "_k").c_str(), k);
}

static void build_params(NAMGene& ng, FeatureCollection& fc, BlobWriter& bw) {
    const int S = ng.statecount, C = Constant::decomp_num_steps;
    bw.scalar_i("statecount", S);
    bw.scalar_i("num_gc_classes", C);
    bw.scalar_i("synchstate", Properties::getIntProperty("/NAMGene/SynchState"));
    bw.scalar_i("utr_option_on", Constant::utr_option_on);
    bw.scalar_i("temperature", (int)Constant::temperature);      // types.cc:443-448: heats the sampling distribution (LLDouble::heated)
    {   // softmasking (extrinsicinfo.cc:1696-1724): every lower-case run becomes a nonexonpart hint of source RM; with the default
        // extrinsic.cfg its bonus (1.15) is the only factor different from 1 — every malus / local malus is 1
        double lb = 0; int plain = 1;
        if (Constant::softmasking) {
            Feature rm(0, 0, nonexonpartF, bothstrands, -1, "RM");
            rm.source = "softmask"; rm.feature = "nep"; rm.score = 0; rm.groupname = ""; rm.priority = -1; rm.mult = 1; rm.gradeclass = 0;
            fc.setBonusMalus(rm);
            lb = log(rm.bonus);
            for (int t = 0; t < NUM_FEATURE_TYPES; t++) if (fc.typeInfo[t].malus != 1.0 || fc.typeInfo[t].localMalus != 1.0) plain = 0;
        }
        bw.scalar_i("softmasking", Constant::softmasking ? 1 : 0);
        bw.scalar_f("softmask_bonus", lb);
        bw.scalar_i("extrinsic_malus_all_one", plain);
    }
    bw.scalar_i("nc_option_on", Constant::nc_option_on);
    bw.scalar_i("dss_gc_allowed", Constant::dss_gc_allowed);
    bw.scalar_i("dss_start", Constant::dss_start); bw.scalar_i("dss_end", Constant::dss_end);
    bw.scalar_i("ass_start", Constant::ass_start); bw.scalar_i("ass_end", Constant::ass_end);
    bw.scalar_i("ass_upwindow_size", Constant::ass_upwindow_size);
    bw.scalar_i("trans_init_window", Constant::trans_init_window);
    bw.scalar_i("init_coding_len", Constant::init_coding_len);
    bw.scalar_i("et_coding_len", Constant::et_coding_len);
    bw.scalar_i("max_exon_len", Constant::max_exon_len);
    bw.scalar_i("min_coding_len", Constant::min_coding_len);
    bw.scalar_i("exon_k", ExonModel::k); bw.scalar_i("intron_k", IntronModel::k); bw.scalar_i("igenic_k", IGenicModel::k);
    bw.scalar_i("min_exon_length", ExonModel::min_exon_length);
    bw.scalar_i("intron_d", IntronModel::d);
    bw.scalar_i("tis_motif_memory", ExonModel::tis_motif_memory);
    bw.scalar_i("transinit_nbins", ExonModel::GCtransInitBinProbs ? ExonModel::GCtransInitBinProbs[0].nbins : 0);
    bw.scalar_f("probNinCoding", log(Constant::probNinCoding));
    bw.scalar_f("ochreprob", Constant::ochreprob > 0 ? log(Constant::ochreprob) : -INFINITY);
    bw.scalar_f("amberprob", Constant::amberprob > 0 ? log(Constant::amberprob) : -INFINITY);
    bw.scalar_f("opalprob", Constant::opalprob > 0 ? log(Constant::opalprob) : -INFINITY);
    {
        std::vector<int32_t> gcw(1);
        try { gcw[0] = Properties::getIntProperty("GCwinsize"); } catch (...) { gcw[0] = 10000; }
        bw.i32("GCwinsize", {1}, gcw);
    }

    std::vector<int32_t> st(S), reach(S);
    for (int i = 0; i < S; i++) { st[i] = (int)ng.stateMap[i]; reach[i] = ng.stateReachable[i]; }
    bw.i32("state_type", {(uint64_t)S}, st);
    bw.i32("state_reachable", {(uint64_t)S}, reach);
    bw.f64("init_probs", {(uint64_t)S}, lgv(ng.initProbs));
    bw.f64("term_probs", {(uint64_t)S}, lgv(ng.termProbs));

    // genetic code: stop codons and start codon probabilities, index = Seq2Int(3) of the codon
    {
        std::vector<int32_t> stop(64); std::vector<double> start(64);
        Seq2Int s2i(3);
        for (int c = 0; c < 64; c++) {
            std::string cod = s2i.inv(c);
            stop[c] = GeneticCode::isStopcodon(cod.c_str());
            start[c] = GeneticCode::start_codons[c] ? lg(GeneticCode::start_codon_probs[c]) : -INFINITY;
        }
        bw.i32("is_stop_codon", {64}, stop);
        bw.f64("start_codon_prob", {64}, start);
    }

    // per-class tables.  Walk the classes exactly as NAMGene::updateToLocalGCEach does (namgene.cc:1564)
    // so that the in-place transition rewrite of IntronModel::updateToLocalGCEach (intronmodel.cc:439-488)
    // is the reference's own arithmetic.
    IntronModel::snippetProbs || (IntronModel::initSnippetProbs(), 0);
    std::vector<double> trans; trans.reserve((size_t)C * S * S);
    std::vector<double> xemi, xinit, xet, iemi, gemi, psi, mal;
    std::vector<std::vector<double>> xpls(ExonModel::k + 1), gpls(IGenicModel::k + 1);
    const char* dummy = "acgtacgtacgtacgtacgtacgtacgtacgt";
    StateModel::sequence = dummy; StateModel::dnalen = strlen(dummy);
    IntronModel::initSnippetProbs();
    static SequenceFeatureCollection nc_sfc(&fc);     // NcModel::initAlgorithms reads hints through StateModel::seqFeatColl (ncmodel.cc:744)
    if (Constant::nc_option_on) {
        nc_sfc.setSeqLen(strlen(dummy));
        if (!StateModel::seqFeatColl) StateModel::setSFC(&nc_sfc);
        NcModel::initSnippetProbs();
    }
    ng.initAlgorithms();
    for (int c = 0; c < C; c++) {
        ng.updateToLocalGCEach(c, 2, 1);
        for (int i = 0; i < S; i++) for (int j = 0; j < S; j++) trans.push_back(lg(ng.transitions[i][j]));
        for (int f = 0; f < 3; f++) { auto t = lgv(ExonModel::emiprobs.probs[f]); xemi.insert(xemi.end(), t.begin(), t.end()); }
        for (int f = 0; f < 3; f++) { auto t = lgv(ExonModel::initemiprobs[f]); xinit.insert(xinit.end(), t.begin(), t.end()); }
        for (int f = 0; f < 3; f++) { auto t = lgv(ExonModel::etemiprobs[f]); xet.insert(xet.end(), t.begin(), t.end()); }
        for (int l = 0; l <= ExonModel::k; l++)
            for (int f = 0; f < 3; f++) { auto t = lgv(ExonModel::Pls[l][f]); xpls[l].insert(xpls[l].end(), t.begin(), t.end()); }
        { auto t = lgv(IntronModel::emiprobs.probs); iemi.insert(iemi.end(), t.begin(), t.end()); }
        { auto t = lgv(IGenicModel::emiprobs.probs); gemi.insert(gemi.end(), t.begin(), t.end()); }
        for (int l = 0; l <= IGenicModel::k; l++) { auto t = lgv(IGenicModel::Pls[l]); gpls[l].insert(gpls[l].end(), t.begin(), t.end()); }
        psi.push_back(lg(IntronModel::probShortIntron)); mal.push_back(IntronModel::mal.doubleValue());
    }
    bw.f64("trans", {(uint64_t)C, (uint64_t)S, (uint64_t)S}, trans);
    uint64_t K1 = 1ull << (2 * (ExonModel::k + 1));
    bw.f64("exon_emi", {(uint64_t)C, 3, K1}, xemi);
    bw.f64("exon_initemi", {(uint64_t)C, 3, K1}, xinit);
    bw.f64("exon_etemi", {(uint64_t)C, 3, K1}, xet);
    for (int l = 0; l <= ExonModel::k; l++) {
        char nm[32]; sprintf(nm, "exon_pls%d", l);
        bw.f64(nm, {(uint64_t)C, 3, 1ull << (2 * (l + 1))}, xpls[l]);
    }
    bw.f64("intron_emi", {(uint64_t)C, 1ull << (2 * (IntronModel::k + 1))}, iemi);
    bw.f64("igenic_emi", {(uint64_t)C, 1ull << (2 * (IGenicModel::k + 1))}, gemi);
    for (int l = 0; l <= IGenicModel::k; l++) {
        char nm[32]; sprintf(nm, "igenic_pls%d", l);
        bw.f64(nm, {(uint64_t)C, 1ull << (2 * (l + 1))}, gpls[l]);
    }
    bw.f64("prob_short_intron", {(uint64_t)C}, psi);
    bw.f64("mal", {(uint64_t)C}, mal);
    export_motif(bw, "tis_motif", ExonModel::GCtransInitMotif, C);
    if (ExonModel::GCtransInitBinProbs && ExonModel::GCtransInitBinProbs[0].nbins >= 1) {
        // TRANSINITBIN (exonmodel.cc:706-713): the start codon x TIS motif probability is mapped to the probability of its bin
        // (BinnedMMGroup::getIndex, merkmal.cc:155-168): per class nbins-1 boundaries and nbins bin probabilities
        const int nb = ExonModel::GCtransInitBinProbs[0].nbins;
        std::vector<double> bb, av;
        for (int c = 0; c < C; c++) {
            BinnedMMGroup& g = ExonModel::GCtransInitBinProbs[c];
            if (g.nbins != nb || (int)g.bb.size() < nb - 1 || (int)g.avprobs.size() < nb) throw ProjectError("augb200 export: TRANSINITBIN tables differ across GC classes");
            for (int i = 0; i < nb - 1; i++) bb.push_back(lg(g.bb[i]));
            for (int i = 0; i < nb; i++) av.push_back(lg(g.avprobs[i]));
        }
        if (nb > 1) bw.f64("tis_bin_bounds", {(uint64_t)C, (uint64_t)(nb - 1)}, bb);
        bw.f64("tis_bin_probs", {(uint64_t)C, (uint64_t)nb}, av);
    }
    export_motif(bw, "ass_motif", IntronModel::GCassMotif, C);

    if (Constant::utr_option_on) {
        // UtrModel tables (utrmodel.cc:540-696), the content tables AFTER the mixing with the intron table (:680-688)
        std::vector<double> u5i, u5, u3, tup;
        for (int c = 0; c < C; c++) {
            { auto t = lgv(UtrModel::GCutr5init_emiprobs[c].probs); u5i.insert(u5i.end(), t.begin(), t.end()); }
            { auto t = lgv(UtrModel::GCutr5_emiprobs[c].probs); u5.insert(u5.end(), t.begin(), t.end()); }
            { auto t = lgv(UtrModel::GCutr3_emiprobs[c].probs); u3.insert(u3.end(), t.begin(), t.end()); }
            { auto t = lgv(UtrModel::GCtssup_emiprobs[c]); tup.insert(tup.end(), t.begin(), t.end()); }
        }
        uint64_t KU = 1ull << (2 * (UtrModel::k + 1));
        bw.scalar_i("utr_k", UtrModel::k); bw.scalar_i("tssup_k", UtrModel::tssup_k);
        bw.f64("utr5init_emi", {(uint64_t)C, KU}, u5i); bw.f64("utr5_emi", {(uint64_t)C, KU}, u5); bw.f64("utr3_emi", {(uint64_t)C, KU}, u3);
        bw.f64("tssup_emi", {(uint64_t)C, 1ull << (2 * (UtrModel::tssup_k + 1))}, tup);
        export_motif(bw, "tss_motif", UtrModel::GCtssMotif, C);
        export_motif(bw, "tsstata_motif", UtrModel::GCtssMotifTATA, C);
        export_motif(bw, "tata_motif", UtrModel::GCtataMotif, C);
        export_motif(bw, "tts_motif", UtrModel::GCttsMotif, C);
        bw.f64("aataaa_probs", {(uint64_t)UtrModel::aataaa_probs.size()}, lgv(UtrModel::aataaa_probs));
        bw.f64("lendist_utr5single", {(uint64_t)UtrModel::lenDist5Single.size()}, lgv(UtrModel::lenDist5Single));
        bw.f64("lendist_utr5initial", {(uint64_t)UtrModel::lenDist5Initial.size()}, lgv(UtrModel::lenDist5Initial));
        bw.f64("lendist_utr5internal", {(uint64_t)UtrModel::lenDist5Internal.size()}, lgv(UtrModel::lenDist5Internal));
        bw.f64("lendist_utr5terminal", {(uint64_t)UtrModel::lenDist5Terminal.size()}, lgv(UtrModel::lenDist5Terminal));
        bw.f64("lendist_utr3single", {(uint64_t)UtrModel::lenDist3Single.size()}, lgv(UtrModel::lenDist3Single));
        bw.f64("lendist_utr3initial", {(uint64_t)UtrModel::lenDist3Initial.size()}, lgv(UtrModel::lenDist3Initial));
        bw.f64("lendist_utr3internal", {(uint64_t)UtrModel::lenDist3Internal.size()}, lgv(UtrModel::lenDist3Internal));
        bw.f64("lendist_utr3terminal", {(uint64_t)UtrModel::lenDist3Terminal.size()}, lgv(UtrModel::lenDist3Terminal));
        bw.f64("taillendist_utr5single", {(uint64_t)UtrModel::tailLenDist5Single.size()}, lgv(UtrModel::tailLenDist5Single));
        bw.f64("taillendist_utr3single", {(uint64_t)UtrModel::tailLenDist3Single.size()}, lgv(UtrModel::tailLenDist3Single));
        bw.scalar_i("utr_max_exon_length", UtrModel::max_exon_length);
        bw.scalar_i("utr_max3singlelength", UtrModel::max3singlelength);
        bw.scalar_i("utr_max3termlength", UtrModel::max3termlength);
        bw.scalar_i("tss_start", UtrModel::tss_start); bw.scalar_i("tss_end", UtrModel::tss_end);
        bw.scalar_i("tata_start", UtrModel::tata_start); bw.scalar_i("tata_end", UtrModel::tata_end);
        bw.scalar_i("d_tss_tata_min", UtrModel::d_tss_tata_min); bw.scalar_i("d_tss_tata_max", UtrModel::d_tss_tata_max);
        bw.scalar_i("tss_upwindow_size", Constant::tss_upwindow_size);
        bw.scalar_i("d_polyasig_cleavage", Constant::d_polyasig_cleavage);
        bw.scalar_i("aataaa_boxlen", UtrModel::aataaa_boxlen);
        bw.scalar_i("tts_spacing", UtrModel::ttsSpacing);
        {   // utrmodel.cc:1850,1872-1878: aataaa_probs[..] * prob_polya, and (1 - prob_polya) * 4^-boxlen without a box
            Double randProb = 1.0 / POWER4TOTHE(UtrModel::aataaa_boxlen);
            Double pp = UtrModel::prob_polya, np = (1 - UtrModel::prob_polya) * randProb;
            bw.scalar_f("log_prob_polya", lg(pp)); bw.scalar_f("log_no_polya", lg(np));
        }
        std::vector<int32_t> isstart(64);
        for (int c = 0; c < 64; c++) isstart[c] = GeneticCode::start_codons[c];
        bw.i32("is_start_codon", {64}, isstart);
    }

    // global tables
    bw.f64("lendist_single", {(uint64_t)ExonModel::lenDistSingle.size()}, lgv(ExonModel::lenDistSingle));
    bw.f64("lendist_initial", {(uint64_t)ExonModel::lenDistInitial.size()}, lgv(ExonModel::lenDistInitial));
    bw.f64("lendist_internal", {(uint64_t)ExonModel::lenDistInternal.size()}, lgv(ExonModel::lenDistInternal));
    bw.f64("lendist_terminal", {(uint64_t)ExonModel::lenDistTerminal.size()}, lgv(ExonModel::lenDistTerminal));
    bw.f64("lendist_intron", {(uint64_t)IntronModel::lenDist.size()}, lgv(IntronModel::lenDist));

    // splice-site pattern scores: the pattern probability AFTER the optional bin mapping
    // (intronmodel.cc:1168-1177 and :1232-1239) — a pure function of the pattern, so tabulated here
    // with the reference's own BinnedMMGroup::getIndex (merkmal.cc:155-168).
    {
        size_t na = IntronModel::assprobs.size(), nd = IntronModel::dssprobs.size();
        std::vector<double> a(na), a2(na), dd(nd), d2(nd);
        for (size_t i = 0; i < na; i++) {
            Double p = IntronModel::assprobs[i], q = p; q *= IntronModel::non_ag_ass_prob;
            if (IntronModel::assBinProbs.nbins >= 1) {
                p = IntronModel::assBinProbs.avprobs[IntronModel::assBinProbs.getIndex(p)];
                q = IntronModel::assBinProbs.avprobs[IntronModel::assBinProbs.getIndex(q)];
            }
            a[i] = lg(p); a2[i] = lg(q);
        }
        for (size_t i = 0; i < nd; i++) {
            Double p = IntronModel::dssprobs[i], q = p; q *= IntronModel::non_gt_dss_prob;
            if (IntronModel::dssBinProbs.nbins >= 1) {
                p = IntronModel::dssBinProbs.avprobs[IntronModel::dssBinProbs.getIndex(p)];
                q = IntronModel::dssBinProbs.avprobs[IntronModel::dssBinProbs.getIndex(q)];
            }
            dd[i] = lg(p); d2[i] = lg(q);
        }
        bw.f64("ass_pattern", {(uint64_t)na}, a); bw.f64("ass_pattern_nonag", {(uint64_t)na}, a2);
        bw.f64("dss_pattern", {(uint64_t)nd}, dd); bw.f64("dss_pattern_nongt", {(uint64_t)nd}, d2);
    }
    // GC-class centroids for ContentStairs (motif.cc:466-505) — exported so the host side can
    // restate computeStairs without the reference: per class (ra, rc, rg, rt), plus the weight matrix.
    {
        ContentDecomposition cd;
        std::vector<double> z;
        for (int i = 0; i < cd.n; i++) { z.push_back(cd.zus[i].ra); z.push_back(cd.zus[i].rc); z.push_back(cd.zus[i].rg); z.push_back(cd.zus[i].rt); }
        bw.f64("gc_centroids", {(uint64_t)cd.n, 4}, z);
        bw.scalar_i("basecount_weighing_type", (int)BaseCount::weithType + 1);
        std::vector<double> wm;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) wm.push_back(BaseCount::weithType == multiNormalKernel ? BaseCount::weighingMatrix[i][j] : (double)(i == j));
        bw.f64("basecount_weight_matrix", {4, 4}, wm);
    }
}
static void export_params(NAMGene& ng, const char* fn, FeatureCollection& fc) {
    BlobWriter bw;
    build_params(ng, fc, bw);
    bw.write(fn);
    fprintf(stderr, "augdump: wrote %zu arrays to %s\n", bw.ents.size(), fn);
}

