/*
 * augshim.cc — the host-side binding of the product (built only where /root/reference exists, into oracle/_ref/).
 *
 * It lives in host/, not in oracle/: this is the code a maintainer adds to the reference, i.e. integration code of the product,
 * while oracle/ holds the checker only.  The reference-side binding of INTEGRATION.md as working code: strong definitions of the three DP
 * entry points of NAMGene,
 *     NAMGene::viterbiAndForward   (reference src/namgene.cc:168-365)
 *     NAMGene::getViterbiPath      (src/namgene.cc:432-510)
 *     NAMGene::getSampledPath      (src/namgene.cc:367-426)
 * that call the C ABI of libaugb200.so (include/augb200.h) instead of running the CPU matrices.
 * oracle/Makefile links them with the UNMODIFIED reference objects (main = src/augustus.cc) after
 * weakening those three symbols in a copy of namgene.o (objcopy --weaken-symbol), so every call site
 * in the reference (findGenes, namgene.cc:776,790,848) lands here.  The result, oracle/_ref/augustus_b200,
 * takes the augustus command line and config/species files and prints GFF through the reference's own
 * post-processing; tests/test_dropin.py compares that GFF with the one of oracle/_ref/augustus.
 *
 * Sequences longer than maxDNAPieceSize work as in the reference: doViterbiPiecewise (namgene.cc:516-676) and its cut search
 * (getNextCutEndPoint, :973-1133) reach the DP only through these three entry points; at a cut the host swaps initProbs / termProbs
 * for the synch-state vectors (:594-603), so the shim keeps one library model per (initial, terminal) vector pair it meets (at most
 * four).  The cut search decodes without sampling; the sampled paths are therefore drawn at the first getSampledPath call, which
 * keeps the rand() stream position of the process exact.
 *
 * No CPU fallback: configurations the library does not decode (hints files, protein profiles, overlap
 * mode, CRF training) raise ProjectError.
 */
#include <map>
#include "aug_export.h"
#include "augb200.h"

namespace {

struct ShimState {
    augb200_model* model = nullptr;                          /* the model for the current initProbs / termProbs */
    std::map<std::string, augb200_model*> models;            /* one per (initProbs, termProbs) pair met so far */
    std::vector<char> blob;
    bool sampled = false;                                    /* samples of the current window drawn */
    augb200_path vit;
    std::vector<augb200_path> samples;
    int next_sample = 0;
    long dnalen = 0;
    bool have = false;
    std::string masked;           /* the window with soft-masked runs in lower case, the rest upper case */
    long calls = 0;
    uint64_t rand_pos = 0;        /* rand() values the sampled paths of this process have consumed so far (vitmatrix.cc:300) */
};
ShimState g;

void fail(const std::string& what, int rc) {
    throw ProjectError("augb200: " + what + ": " + augb200_strerror(rc) + " " + augb200_last_cuda_error());
}

/* key of a model variant: the bytes of the initial and terminal vectors as NAMGene holds them right now */
std::string probs_key(const std::vector<Double>& a, const std::vector<Double>& b) {
    std::string k;
    for (const std::vector<Double>* v : {&a, &b})
        for (size_t i = 0; i < v->size(); i++) { double x = (*v)[i].log(); k.append((const char*)&x, sizeof x); }
    return k;
}

void ensure_model(NAMGene& ng, const std::string& key) {
    std::map<std::string, augb200_model*>::iterator it = g.models.find(key);
    if (it != g.models.end()) { g.model = it->second; return; }
    /* the export reads extrinsic.cfg and swaps GC classes through the models; both chat on cout / cerr, which must not leak into
     * the GFF (the front end already printed its own "Sources of extrinsic information" line) */
    std::ostringstream sink;
    std::streambuf* oc = std::cout.rdbuf(sink.rdbuf()); std::streambuf* oe = std::cerr.rdbuf(sink.rdbuf());
    try {
        FeatureCollection fc;
        if (Constant::softmasking) fc.readExtrinsicCFGFile();
        BlobWriter bw;
        build_params(ng, fc, bw);                            /* init_probs / term_probs = the vectors of this piece */
        g.blob = bw.bytes();
    } catch (...) { std::cout.rdbuf(oc); std::cerr.rdbuf(oe); throw; }
    std::cout.rdbuf(oc); std::cerr.rdbuf(oe);
    int dev = 0;
    if (const char* e = getenv("AUGB200_DEVICE")) dev = atoi(e);
    augb200_model* m = nullptr;
    int rc = augb200_model_create(g.blob.data(), g.blob.size(), dev, &m);
    if (rc) fail("model_create", rc);
    g.models[key] = g.model = m;
    if (getenv("AUGSHIM_VERBOSE")) fprintf(stderr, "augshim: model %zu with %d states, %d GC classes on device %d\n", g.models.size(),
                                            augb200_model_statecount(g.model), augb200_model_num_gc_classes(g.model), dev);
}

StatePath* to_statepath(const augb200_path& p, const char* seqname) {
    if (p.status == AUGB200_ERR_NO_PATH) throw ProjectError("No feasible path found in HMM");                       /* namgene.cc:455-457 */
    if (p.status) throw ProjectError(std::string("Viterbi got stuck: ") + augb200_strerror(p.status));             /* :493-496 */
    StatePath* sp = new StatePath();
    if (seqname) sp->seqname = seqname;
    sp->pathemiProb = LLDouble::exp(p.log_prob);
    for (int i = p.n - 1; i >= 0; i--) {             /* push() prepends (gene.hh:217-220): go right to left */
        State* s = new State(p.begin[i], p.end[i], (StateType)p.type[i]);
        s->truncated = (char)p.truncated[i];
        sp->push(s);
    }
    return sp;
}

}  // namespace

void NAMGene::viterbiAndForward(const char* dna, bool useProfile) {
    if (useProfile || profileModel) throw ProjectError("augb200: protein profile models are not decoded on the GPU");
    if (Constant::overlapmode) throw ProjectError("augb200: overlap mode is not decoded on the GPU");
    if (inCRFTraining) throw ProjectError("augb200: CRF training runs on the CPU build");
    {   /* --emiprobs re-scores the predicted path against the CPU matrices (getPathEmiProb, augustus.cc:423-441): not available here */
        bool emi = false;
        try { emi = Properties::getBoolProperty("emiprobs"); } catch (...) {}
        if (emi) throw ProjectError("augb200: --emiprobs needs the DP matrices of the CPU build");
    }
    const long n = (long)strlen(dna);
    /* the DP sees a lower-cased sequence (SequenceFeatureCollection::prepare, extrinsicinfo.cc:1726-1727); soft-masked runs
     * arrive as nonexonpart hints of source RM (:1696-1724).  The library takes the case of the window instead, so rebuild it. */
    g.masked.assign(dna, dna + n);
    for (long i = 0; i < n; i++) g.masked[i] = (char)toupper((unsigned char)g.masked[i]);
    SequenceFeatureCollection* sfc = StateModel::seqFeatColl;
    if (sfc) {
        for (int t = 0; t < NUM_FEATURE_TYPES; t++) {
            const std::list<Feature>& fl = sfc->featureLists[t];
            for (std::list<Feature>::const_iterator it = fl.begin(); it != fl.end(); ++it) {
                const bool rm = t == (int)nonexonpartF && it->source == "softmask";
                if (!rm) throw ProjectError("augb200: hints other than softmasking are not decoded on the GPU (feature type " + std::to_string(t) + " source " + it->source + ")");
                if (!it->active) continue;
                for (long p = std::max<long>(0, it->start); p <= it->end && p < n; p++) g.masked[p] = (char)tolower((unsigned char)g.masked[p]);
            }
        }
    }
    ensure_model(*this, probs_key(initProbs, termProbs));
    augb200_window w; w.dna = g.masked.data(); w.length = (int32_t)n; w.gc_class = nullptr;
    /* cs is public state other reference code reads (printing of GC classes); keep it filled as the original does (:228) */
    cs.computeStairs(dna);
    w.gc_class = cs.idx;
    /* Viterbi only: whether sampled paths are wanted is known at the first getSampledPath call (findGenes samples, the cut search
     * of getNextCutEndPoint does not, and only sampling moves the rand() stream) */
    g.samples.clear(); g.sampled = false;
    int rc = augb200_decode(g.model, &w, &g.vit);
    if (rc) fail("decode", rc);
    g.next_sample = 0; g.dnalen = n; g.have = true; g.calls++;
}

StatePath* NAMGene::getViterbiPath(const char* dna, const char* seqname) {
    if (!g.have) throw ProjectError("augb200: getViterbiPath before viterbiAndForward");
    return to_statepath(g.vit, seqname);
}

StatePath* NAMGene::getSampledPath(const char* dna, const char* seqname) {
    if (!needForwardTable) return new StatePath();                                   /* namgene.cc:369-370 */
    if (!g.have) throw ProjectError("augb200: getSampledPath before viterbiAndForward");
    if (!g.sampled) {
        /* forward matrix + the sampleiterations-1 paths findGenes will ask for (namgene.cc:848), from the process-wide stream position */
        if (sampleiterations < 2) throw ProjectError("augb200: sampled path requested with --sample < 2");
        augb200_window w; w.dna = g.masked.data(); w.length = (int32_t)g.dnalen; w.gc_class = cs.idx;
        g.samples.assign(sampleiterations - 1, augb200_path());
        int rc;
        if ((rc = augb200_set_rand_position(g.model, g.rand_pos))) fail("set_rand_position", rc);
        if ((rc = augb200_decode_batch_sampling(g.model, 1, &w, sampleiterations, &g.vit, g.samples.data()))) fail("decode_batch_sampling", rc);
        g.rand_pos += (uint64_t)augb200_last_rand_consumed(g.model);
        g.sampled = true;
    }
    if (g.next_sample >= (int)g.samples.size()) throw ProjectError("augb200: more sampled paths requested than --sample");
    return to_statepath(g.samples[g.next_sample++], seqname);
}
