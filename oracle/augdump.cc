/*
 * augdump.cc — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Our own driver, linked against the UNMODIFIED reference objects built by oracle/Makefile.  It
 *   (1) exports the parameter blob (include/augb200_params.h) from the reference's static tables
 *       after StateModel::readAllParameters() (reference src/augustus.cc:176), and
 *   (2) runs the reference DP (NAMGene::viterbiAndForward, namgene.cc:168; getViterbiPath, :432;
 *       getSampledPath, :367) on each input sequence and dumps the state path, the GC-class stairs
 *       and optionally the full Viterbi / forward matrices, so that the CPU restatement
 *       (oracle/ghmm_oracle.c) and the CUDA path can be checked cell-by-cell, not only end-to-end.
 *
 * Control via environment (the command line is the normal augustus one, because
 * Properties::init validates it against config/parameters/aug_cmdln_parameters.json):
 *   AUGDUMP_PARAMS=<file>   write the parameter blob
 *   AUGDUMP_PATH=<file>     write paths (text)
 *   AUGDUMP_MATRIX=<prefix> write <prefix>.<seqno>.vit [and .fwd]  (L x S doubles, log space, -inf = absent)
 *
 * Access to the reference's private statics: this translation unit (and only this one) is compiled
 * with private/protected re-defined AFTER all standard headers are included; class layout is
 * unaffected by access specifiers under the Itanium ABI.
 */
#include "aug_export.h"

/* ---------------------------------------------------------------- main */
int main(int argc, char* argv[]) {
    const char* fparams = getenv("AUGDUMP_PARAMS");
    const char* fpath = getenv("AUGDUMP_PATH");
    const char* fmat = getenv("AUGDUMP_MATRIX");
    LLDouble::setOutputPrecision(3);
    try {
        Properties::init(argc, argv);
        Constant::init();
        Gene::init();
        GeneticCode::init();
        StateModel::init();
        string filename = Properties::getProperty(INPUTFILE_KEY);
        GBProcessor gbank(filename);
        FeatureCollection extrinsicFeatures;
        if (Constant::softmasking) extrinsicFeatures.readExtrinsicCFGFile();
        BaseCount::init();
        PP::initConstants();
        NAMGene namgene;
        StateModel::readAllParameters();
        if (fparams) { export_params(namgene, fparams, extrinsicFeatures); }
        if (!fpath && !fmat) return 0;

        FILE* fp = fpath ? fopen(fpath, "w") : NULL;
        AnnoSequence* seq = gbank.getSequenceList();
        int seqno = 0;
        const int S = namgene.statecount;
        while (seq) {
            AnnoSequence* cur = seq; seq = seq->next; cur->next = NULL; seqno++;
            SequenceFeatureCollection& sfc = extrinsicFeatures.getSequenceFeatureCollection(cur->seqname);
            sfc.prepare(cur, false);
            const char* dna = cur->sequence;
            long n = strlen(dna);
            // as NAMGene::getStepGenes (namgene.cc:695-700)
            sfc.setSeqLen(n); sfc.computeHintedSites(dna); sfc.makeGroups(); sfc.prepareLocalMalus(dna);
            StateModel::setSFC(&sfc);
            if (namgene.sampleiterations < 1) namgene.sampleiterations = 1;
            namgene.viterbiAndForward(dna, false);
            if (fmat) {
                char fn[4096]; snprintf(fn, sizeof fn, "%s.%d.vit", fmat, seqno);
                FILE* f = fopen(fn, "wb");
                std::vector<double> col(S);
                for (long j = 0; j < n; j++) {
                    for (int s = 0; s < S; s++) col[s] = lg(namgene.viterbi[j].get(s));
                    fwrite(col.data(), 8, S, f);
                }
                fclose(f);
                if (namgene.needForwardTable) {
                    snprintf(fn, sizeof fn, "%s.%d.fwd", fmat, seqno);
                    f = fopen(fn, "wb");
                    for (long j = 0; j < n; j++) {
                        for (int s = 0; s < S; s++) col[s] = lg(namgene.forward[j].get(s));
                        fwrite(col.data(), 8, S, f);
                    }
                    fclose(f);
                }
            }
            if (fp) {
                StatePath* p = namgene.getViterbiPath(dna, "");
                int cnt = 0; for (State* s = p->first; s; s = s->next) cnt++;
                fprintf(fp, "seq %s %ld %d %.17g\n", cur->seqname, n, cnt, lg(p->pathemiProb));
                // GC stairs, run-length
                int last = -1;
                for (long j = 0; j < n; j++) if (namgene.cs.idx[j] != last) { last = namgene.cs.idx[j]; fprintf(fp, "gc %ld %d\n", j, last); }
                for (State* s = p->first; s; s = s->next)
                    fprintf(fp, "state %d %ld %ld %d\n", (int)s->type, (long)s->begin, (long)s->end, (int)s->truncated);
                delete p;
                if (namgene.needForwardTable && namgene.sampleiterations > 1) {
                    // as NAMGene::findGenes (namgene.cc:807-808): the Viterbi matrix is cleared before sampling
                    namgene.viterbi.assign(n, S);
                    for (int it = 0; it < namgene.sampleiterations - 1; it++) {
                        StatePath* sp = namgene.getSampledPath(dna, "");
                        int c2 = 0; for (State* s = sp->first; s; s = s->next) c2++;
                        fprintf(fp, "sample %d %d %.17g\n", it, c2, lg(sp->pathemiProb));
                        for (State* s = sp->first; s; s = s->next)
                            fprintf(fp, "sstate %d %ld %ld %d\n", (int)s->type, (long)s->begin, (long)s->end, (int)s->truncated);
                        delete sp;
                    }
                }
            }
            cur->next = seq;
        }
        if (fp) fclose(fp);
    } catch (ProjectError& err) {
        cerr << "augdump: ERROR " << err.getMessage() << endl;
        return 1;
    }
    return 0;
}
