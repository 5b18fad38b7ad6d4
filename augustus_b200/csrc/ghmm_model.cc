/* ghmm_model.cc — see ghmm_model.h */
#include "ghmm_model.h"

#include <cstdio>

namespace augb {

namespace {
struct Reader {
    HostModel::BlobView b; std::string* err; bool ok = true;
    /* an entry whose payload lies inside the blob and holds at least `min_bytes` of the expected type */
    const augb200_blob_entry* need(const char* name, uint32_t dtype, size_t min_bytes) {
        const augb200_blob_entry* e = b.find(name);
        if (!e) { ok = false; *err = std::string("blob entry missing: ") + name; return nullptr; }
        if (e->offset > b.n || e->nbytes > b.n - e->offset) { ok = false; *err = std::string("blob entry out of range: ") + name; return nullptr; }
        if (e->dtype != dtype || e->nbytes < min_bytes || (e->offset & 7)) { ok = false; *err = std::string("blob entry of unexpected type / size: ") + name; return nullptr; }
        return e;
    }
    int i32(const char* name) { auto e = need(name, AUGB200_DT_I32, 4); int32_t v = 0; if (e) memcpy(&v, b.p + e->offset, 4); return v; }
    double f64(const char* name) { auto e = need(name, AUGB200_DT_F64, 8); double v = 0; if (e) memcpy(&v, b.p + e->offset, 8); return v; }
    const double* darr(const char* name, size_t* n) { auto e = need(name, AUGB200_DT_F64, 0); if (!e) { *n = 0; return nullptr; } *n = e->nbytes / 8; return (const double*)(b.p + e->offset); }
    const int32_t* iarr(const char* name, size_t* n) { auto e = need(name, AUGB200_DT_I32, 0); if (!e) { *n = 0; return nullptr; } *n = e->nbytes / 4; return (const int32_t*)(b.p + e->offset); }
};
}  // namespace

static void classify(StateDesc& s, const DevModel& m) {
    int t = s.type;
    s.fwd = t < 36; s.frame = 0; s.chain = -1; s.feeds = -1; s.kind = -1; s.ek = -1;
    s.beginPartLen = s.innerPartOffset = s.baseOffset = s.innerPartEndOffset = 0;
    if (t == T_IGENIC) { s.kind = K_IGENIC; s.fwd = 1; return; }
    int e = -1;
    if (t == T_SINGLE) e = E_SINGLE;
    else if (t >= T_INITIAL0 && t < T_INITIAL0 + 3) { e = E_INITIAL; s.frame = t - T_INITIAL0; }
    else if (t >= T_INTERNAL0 && t < T_INTERNAL0 + 3) { e = E_INTERNAL; s.frame = t - T_INTERNAL0; }
    else if (t == T_TERMINAL) e = E_TERMINAL;
    else if (t == T_RSINGLE) { e = E_RSINGLE; s.frame = 2; }
    else if (t == T_RINITIAL) { e = E_RINITIAL; s.frame = 2; }
    else if (t >= T_RINTERNAL0 && t < T_RINTERNAL0 + 3) { e = E_RINTERNAL; s.frame = t - T_RINTERNAL0; }
    else if (t >= T_RTERMINAL0 && t < T_RTERMINAL0 + 3) { e = E_RTERMINAL; s.frame = t - T_RTERMINAL0; }
    if (e >= 0) {
        s.kind = K_EXON; s.ek = e;
        switch (e) {   /* ExonModel::ExonModel, exonmodel.cc:231-279 */
        case E_SINGLE: case E_INITIAL: s.beginPartLen = 3 + m.tiw; s.innerPartOffset = 3; break;
        case E_RSINGLE: case E_RTERMINAL: s.beginPartLen = s.innerPartOffset = 3; break;
        default: s.beginPartLen = 0; s.innerPartOffset = s.fwd ? m.ass_end : m.dss_start;
        }
        if (e == E_SINGLE || e == E_TERMINAL) { s.baseOffset = 0; s.innerPartEndOffset = 3; }
        else if (e == E_RSINGLE || e == E_RINITIAL) { s.baseOffset = -m.tiw; s.innerPartEndOffset = 3; }
        else { s.baseOffset = s.innerPartEndOffset = s.fwd ? m.dss_start : m.ass_end; }
        return;
    }
    s.uk = -1; s.u5 = 0;
    if (t >= T_NCSINGLE && t < T_NCSINGLE + 12) {
        /* NcModel (ncmodel.cc).  Without hints a transcript boundary of a non-coding gene has probability zero (precomputeTxEndProbs,
         * :744-826), so only the one-base intron and the splice-site-to-splice-site internal exon ever hold cells beyond column 0 */
        if (!m.utr || !m.nc) return;
        s.fwd = t < T_RNCSINGLE; s.uk = (int8_t)((t - T_NCSINGLE) % 6); s.u5 = 2;
        s.kind = (s.uk == U_INTRON || s.uk == U_INTERNAL) ? K_UTR : K_DEAD;
        return;
    }
    if ((t >= T_UTR5SINGLE && t <= T_UTR3TERM) || (t >= T_RUTR5SINGLE && t <= T_RUTR3TERM)) {
        if (!m.utr) return;
        int o = t - (s.fwd ? T_UTR5SINGLE : T_RUTR5SINGLE);
        s.kind = K_UTR; s.u5 = o < 6; s.uk = (int8_t)(o % 6);
        return;
    }
    int base = s.fwd ? T_LESSD0 : T_RLESSD0, off = t - base;
    if (off < 0 || off >= 15) return;     /* kind stays -1: unsupported */
    s.frame = off / 5;
    static const int kinds[5] = {K_LESSD, K_LONGDSS, K_EQUALD, K_GEO, K_LONGASS};
    s.kind = kinds[off % 5];
}

int HostModel::build(const void* blob, size_t nbytes) {
    if (nbytes < sizeof(augb200_blob_header) || memcmp(blob, AUGB200_BLOB_MAGIC, 8)) { err = "bad magic"; return AUGB200_ERR_BAD_BLOB; }
    const augb200_blob_header* h = (const augb200_blob_header*)blob;
    if (h->version != AUGB200_BLOB_VERSION || sizeof *h + (size_t)h->n_entries * sizeof(augb200_blob_entry) > nbytes) { err = "bad header"; return AUGB200_ERR_BAD_BLOB; }
    Reader r{{(const char*)blob, nbytes}, &err};
    memset(&dm, 0, sizeof dm);
    DevModel& m = dm;
    m.S = r.i32("statecount"); m.C = r.i32("num_gc_classes");
    m.k = r.i32("exon_k");
    int ik = r.i32("intron_k"), gk = r.i32("igenic_k");
    m.d = r.i32("intron_d");
    m.dss_start = r.i32("dss_start"); m.dss_end = r.i32("dss_end"); m.ass_start = r.i32("ass_start"); m.ass_end = r.i32("ass_end");
    m.ass_up = r.i32("ass_upwindow_size"); m.tiw = r.i32("trans_init_window");
    m.init_len = r.i32("init_coding_len"); m.et_len = r.i32("et_coding_len");
    m.max_exon_len = r.i32("max_exon_len"); m.min_exon_length = r.i32("min_exon_length");
    m.dss_gc_allowed = r.i32("dss_gc_allowed"); m.GCwinsize = r.i32("GCwinsize"); m.weighing = r.i32("basecount_weighing_type");
    m.tis_n = r.i32("tis_motif_n"); m.tis_k = r.i32("tis_motif_k"); m.assm_n = r.i32("ass_motif_n"); m.assm_k = r.i32("ass_motif_k");
    int nbins = r.i32("transinit_nbins"), utr = r.i32("utr_option_on"), nc = r.i32("nc_option_on");
    if (!r.ok) return AUGB200_ERR_BAD_BLOB;
    if (m.S < 1 || m.S > MAXS || m.C < 1 || m.C > MAXC) { err = "state / class count out of range"; return AUGB200_ERR_UNSUPPORTED; }
    {   /* geometry integers feed shift counts, table sizes and loop bounds: check them before anything is computed from them */
        auto in = [](int v, int lo, int hi) { return v >= lo && v <= hi; };
        if (!in(m.dss_start, 0, 12) || !in(m.dss_end, 0, 12) || !in(m.ass_start, 0, 12) || !in(m.ass_end, 0, 12) || m.dss_start + m.dss_end > 14 || m.ass_start + m.ass_end > 14 ||
            !in(m.ass_up, 0, 400) || !in(m.tiw, 0, 400) || !in(m.init_len, 0, 1000) || !in(m.et_len, 0, 1000) || !in(m.d, 1, 1 << 20) ||
            !in(m.max_exon_len, 1, 1 << 24) || !in(m.min_exon_length, 0, 1 << 20) || !in(m.GCwinsize, 0, 1 << 28) || !in(m.weighing, 0, 16) ||
            !in(m.tis_n, 0, 400) || !in(m.tis_k, 0, 4) || !in(m.assm_n, 0, 400) || !in(m.assm_k, 0, 4) || !in(nbins, -1, 64))
            { err = "model geometry out of range"; return AUGB200_ERR_BAD_BLOB; }
    }
    if (ik != m.k || gk != m.k || m.k < 1 || m.k > 4) { err = "content model orders must be equal and <= 4"; return AUGB200_ERR_UNSUPPORTED; }
    m.tis_nbins = nbins > 0 ? nbins : 0;
    m.utr = utr ? 1 : 0; m.nc = nc ? 1 : 0;
    if (nc && !utr) { err = "nc states need the UTR states"; return AUGB200_ERR_UNSUPPORTED; }
    if (r.b.find("softmasking")) {          /* blobs written before softmasking support carry no entry: off */
        m.softmask = r.i32("softmasking") ? 1 : 0; m.nep_bonus = quantize(r.f64("softmask_bonus"));
        if (m.softmask && !r.i32("extrinsic_malus_all_one")) { err = "extrinsic configurations with a malus are not supported"; return AUGB200_ERR_UNSUPPORTED; }
    }
    if (m.nc && m.softmask) { err = "nc states with softmasking are not supported yet"; return AUGB200_ERR_UNSUPPORTED; }
    /* --temperature heats the forward summands of the sampling pass (LLDouble::heated, lldouble.cc:209-264): exponent (8 - T) / 8 in the log domain */
    m.heat = 1.0;
    if (r.b.find("temperature")) {
        const int T = r.i32("temperature");
        if (T < 0 || T > 7) { err = "sampling temperature must be 0 .. 7"; return AUGB200_ERR_UNSUPPORTED; }
        m.heat = (8.0 - T) / 8.0;
    }
    m.dStateLen = m.d - 2 - m.dss_end - m.ass_start - 2 - m.ass_up;     /* intronmodel.cc:519-520 */
    if (m.dStateLen < 1) { err = "d too small"; return AUGB200_ERR_UNSUPPORTED; }
    /* columns around a GC-class boundary in which the SnippetProbs memo is restated (Sweep::snip_get): dStateLen columns before it every
     * key a later request can reach is created, 2 * dStateLen after it every piece in play carries the new class */
    m.snip_before = m.dStateLen + 8; m.snip_after = 2 * m.dStateLen + 16;
    if (m.C > 1 && m.dStateLen + 16 > SNIP_RING) { err = "intron d too large for the memo ring of multi-class models"; return AUGB200_ERR_UNSUPPORTED; }

    /* ---- tables ---- */
    tab.clear(); off.clear();
    auto push = [&](const char* name, size_t expect) -> size_t {
        size_t n; const double* d = r.darr(name, &n);
        size_t o = tab.size();
        if (!d) return o;
        if (expect && n != expect) { r.ok = false; err = std::string("unexpected size of ") + name; return o; }
        for (size_t i = 0; i < n; i++) tab.push_back(quantize(d[i]));
        return o;
    };
    const size_t K1 = (size_t)1 << (2 * (m.k + 1));
    size_t o_init = push("init_probs", m.S), o_term = push("term_probs", m.S), o_trans = push("trans", (size_t)m.C * m.S * m.S);
    size_t o_xemi = push("exon_emi", m.C * 3 * K1), o_xinit = push("exon_initemi", m.C * 3 * K1), o_xet = push("exon_etemi", m.C * 3 * K1);
    size_t o_xpls[5] = {0, 0, 0, 0, 0};
    for (int l = 0; l <= m.k; l++) { char nm[32]; snprintf(nm, sizeof nm, "exon_pls%d", l); o_xpls[l] = push(nm, (size_t)m.C * 3 << (2 * (l + 1))); }
    size_t o_iemi = push("intron_emi", m.C * K1), o_gemi = push("igenic_emi", m.C * K1);
    size_t o_tis = push("tis_motif", (size_t)m.C * m.tis_n << (2 * (m.tis_k + 1))), o_assm = push("ass_motif", (size_t)m.C * m.assm_n << (2 * (m.assm_k + 1)));
    size_t o_tbb = m.tis_nbins > 1 ? push("tis_bin_bounds", (size_t)m.C * (m.tis_nbins - 1)) : 0, o_tbp = m.tis_nbins > 0 ? push("tis_bin_probs", (size_t)m.C * m.tis_nbins) : 0;
    size_t n0 = tab.size();
    size_t o_lds = push("lendist_single", 0); m.n_ld_exon = (int)(tab.size() - n0);
    size_t o_ldi = push("lendist_initial", m.n_ld_exon), o_ldn = push("lendist_internal", m.n_ld_exon), o_ldt = push("lendist_terminal", m.n_ld_exon);
    n0 = tab.size();
    size_t o_ldx = push("lendist_intron", 0); m.n_ld_intron = (int)(tab.size() - n0);
    size_t npat_a = (size_t)1 << (2 * (m.ass_start + m.ass_end)), npat_d = (size_t)1 << (2 * (m.dss_start + m.dss_end));
    size_t o_ap = push("ass_pattern", npat_a), o_apn = push("ass_pattern_nonag", npat_a), o_dp = push("dss_pattern", npat_d), o_dpn = push("dss_pattern_nongt", npat_d);
    /* ---- UtrModel tables (utrmodel.cc:540-696) ---- */
    size_t o_u5i = 0, o_u5 = 0, o_u3 = 0, o_tup = 0, o_tssm = 0, o_tsstm = 0, o_tatam = 0, o_ttsm = 0, o_aat = 0, o_uld[10] = {0};
    if (m.utr) {
        if (r.i32("utr_k") != m.k) { err = "content model orders must be equal"; return AUGB200_ERR_UNSUPPORTED; }
        m.tssup_k = r.i32("tssup_k"); m.tss_start = r.i32("tss_start"); m.tss_end = r.i32("tss_end");
        m.tata_start = r.i32("tata_start"); m.tata_end = r.i32("tata_end"); m.d_tata_min = r.i32("d_tss_tata_min"); m.d_tata_max = r.i32("d_tss_tata_max");
        m.tuw = r.i32("tss_upwindow_size"); m.dpc = r.i32("d_polyasig_cleavage"); m.boxlen = r.i32("aataaa_boxlen"); m.tts_spacing = r.i32("tts_spacing");
        m.umax = r.i32("utr_max_exon_length"); m.umax3s = r.i32("utr_max3singlelength"); m.umax3t = r.i32("utr_max3termlength");
        m.tssm_n = r.i32("tss_motif_n"); m.tssm_k = r.i32("tss_motif_k"); m.tsstm_n = r.i32("tsstata_motif_n"); m.tsstm_k = r.i32("tsstata_motif_k");
        m.tatam_n = r.i32("tata_motif_n"); m.tatam_k = r.i32("tata_motif_k"); m.ttsm_n = r.i32("tts_motif_n"); m.ttsm_k = r.i32("tts_motif_k");
        if (!r.ok) return AUGB200_ERR_BAD_BLOB;
        if (m.tssup_k < 0 || m.tssup_k > m.k || m.tssm_k > m.k || m.tsstm_k > m.k || m.tatam_k > m.k || m.ttsm_k > m.k || m.boxlen < 1 || m.boxlen > 8 || m.tts_spacing < 1)
            { err = "UTR motif orders / polyA box out of range"; return AUGB200_ERR_UNSUPPORTED; }
        o_u5i = push("utr5init_emi", m.C * K1); o_u5 = push("utr5_emi", m.C * K1); o_u3 = push("utr3_emi", m.C * K1);
        o_tup = push("tssup_emi", (size_t)m.C << (2 * (m.tssup_k + 1)));
        o_tssm = push("tss_motif", (size_t)m.C * m.tssm_n << (2 * (m.tssm_k + 1))); o_tsstm = push("tsstata_motif", (size_t)m.C * m.tsstm_n << (2 * (m.tsstm_k + 1)));
        o_tatam = push("tata_motif", (size_t)m.C * m.tatam_n << (2 * (m.tatam_k + 1))); o_ttsm = push("tts_motif", (size_t)m.C * m.ttsm_n << (2 * (m.ttsm_k + 1)));
        o_aat = push("aataaa_probs", (size_t)1 << (2 * m.boxlen));
        static const char* ldn[10] = {"lendist_utr5single", "lendist_utr5initial", "lendist_utr5internal", "lendist_utr5terminal", "lendist_utr3single",
                                      "lendist_utr3initial", "lendist_utr3internal", "lendist_utr3terminal", "taillendist_utr5single", "taillendist_utr3single"};
        for (int i = 0; i < 10; i++) { size_t b0 = tab.size(); o_uld[i] = push(ldn[i], 0); m.n_uld[i] = (int)(tab.size() - b0); }
        m.log_polya = quantize(r.f64("log_prob_polya")); m.log_nopolya = quantize(r.f64("log_no_polya")); m.log2 = quantize(log(2.0));
        size_t n3; const int32_t* iss = r.iarr("is_start_codon", &n3);
        if (!r.ok || n3 != 64) { err = "bad UTR tables"; return AUGB200_ERR_BAD_BLOB; }
        for (int i = 0; i < 64; i++) m.isstart[i] = (uint8_t)iss[i];
    }
    if (!r.ok) return AUGB200_ERR_BAD_BLOB;
    if (m.n_ld_intron < m.d + 1) { err = "intron length distribution shorter than d"; return AUGB200_ERR_BAD_BLOB; }
    /* igenic emission of the first k columns (igenicmodel.cc:342-354), incl. the normaliser quirk */
    size_t o_gfirst = tab.size();
    {
        const size_t per = ((size_t)1 << (2 * (m.k + 2))) / 3 - 1;     /* sum_{j<=k} 4^(j+1) = (4^(k+2)-4)/3 */
        tab.resize(o_gfirst + per * m.C, SC_NEG);
        for (int j = 0; j <= m.k; j++) {
            char nm[32]; snprintf(nm, sizeof nm, "igenic_pls%d", j);
            size_t n; const double* P = r.darr(nm, &n);
            size_t w = (size_t)1 << (2 * (j + 1));
            if (!P || n != w * m.C) { err = "bad igenic_pls"; return AUGB200_ERR_BAD_BLOB; }
            size_t joff = (w - 4) / 3;
            for (int c = 0; c < m.C; c++)
                for (size_t b = 0; b < w; b++) {
                    const double* Pc = P + c * w;
                    size_t q4 = b / 4;
                    double den = 0; bool okd = true;
                    for (int i = 0; i < 4; i++) { if (q4 + i >= w) { okd = false; break; } den += exp(Pc[q4 + i]); }
                    tab[o_gfirst + c * per + joff + b] = okd ? quantize(Pc[b] - log(den)) : SC_NEG;
                }
        }
    }
    {
        size_t n; const double* sp = r.darr("start_codon_prob", &n); size_t n2; const int32_t* st = r.iarr("is_stop_codon", &n2);
        if (!r.ok || n != 64 || n2 != 64) { err = "bad codon tables"; return AUGB200_ERR_BAD_BLOB; }
        for (int i = 0; i < 64; i++) { m.startp[i] = quantize(sp[i]); m.isstop[i] = (uint8_t)st[i]; }
    }
    m.ochre = quantize(r.f64("ochreprob")); m.amber = quantize(r.f64("amberprob")); m.opal = quantize(r.f64("opalprob"));
    m.probN = quantize(r.f64("probNinCoding")); m.log025 = quantize(log(0.25)); m.log3 = quantize(log(3.0));
    m.ass_invalid_pat = quantize(log(0.001) + (m.ass_start + m.ass_end) * log(0.25));
    {
        size_t n; const double* cen = r.darr("gc_centroids", &n); size_t n2; const double* wmx = r.darr("basecount_weight_matrix", &n2);
        if (!r.ok || n % 4 || n / 4 > MAXC * 16 || n2 != 16) { err = "bad gc tables"; return AUGB200_ERR_BAD_BLOB; }
        m.ncent = (int)(n / 4);
        for (int i = 0; i < m.ncent; i++) for (int j = 0; j < 4; j++) m.centroids[i][j] = cen[4 * i + j];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m.wm[i][j] = wmx[4 * i + j];
    }
    if (!r.ok) return AUGB200_ERR_BAD_BLOB;

    /* ---- states ---- */
    size_t n; const int32_t* stt = r.iarr("state_type", &n); size_t n2; const int32_t* reach = r.iarr("state_reachable", &n2);
    if (!r.ok || (int)n != m.S || (int)n2 != m.S) { err = "bad state tables"; return AUGB200_ERR_BAD_BLOB; }
    for (int c = 0; c < NCHAIN; c++) m.chain_state[c] = -1;
    for (int d = 0; d < 2; d++) for (int f = 0; f < 3; f++) m.r_longdss[d][f] = m.r_lessd[d][f] = m.r_equald[d][f] = m.r_longass[d][f] = -1;
    m.r_single = m.r_terminal = m.r_rsingle = m.r_rinitial = -1;
    for (int f = 0; f < 3; f++) m.r_initial[f] = m.r_internal[f] = m.r_rinternal[f] = m.r_rterminal[f] = -1;
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < 6; c++) m.r_utr[a][b][c] = -1;
    for (int i = 0; i < 18; i++) m.uslot[i] = -1;
    const sc_t* T = tab.data() + o_trans;
    for (int s = 0; s < m.S; s++) {
        StateDesc& sd = m.st[s]; sd.type = (int16_t)stt[s]; classify(sd, m);
        if (sd.kind < 0) { err = "state type " + std::to_string(stt[s]) + " is not supported"; return AUGB200_ERR_UNSUPPORTED; }
        if (!reach[s]) { err = "models with unreachable states are not supported"; return AUGB200_ERR_UNSUPPORTED; }
        sd.nanc = 0;
        for (int a = 0; a < m.S; a++) if (!isneg(T[(size_t)a * m.S + s])) {
            if (sd.nanc >= MAXANC) { err = "too many ancestors"; return AUGB200_ERR_UNSUPPORTED; }
            sd.anc[sd.nanc++] = (int8_t)a;
        }
        for (int c = 1; c < m.C; c++)
            for (int a = 0; a < m.S; a++)
                if (isneg(T[((size_t)c * m.S + a) * m.S + s]) != isneg(T[(size_t)a * m.S + s])) { err = "transition support differs between GC classes"; return AUGB200_ERR_UNSUPPORTED; }
        int dir = sd.fwd ? 0 : 1, f = sd.frame;
        int8_t* slot = nullptr;
        if (sd.u5 == 2) {                      /* nc states: two chains, two table-driven exon states, the rest never evaluated */
            if (sd.uk == U_INTRON) { sd.chain = (int8_t)(CH_NC + dir); slot = &m.chain_state[sd.chain]; }
            else if (sd.uk == U_INTERNAL) { sd.ek = (int8_t)(16 + dir); slot = &m.uslot[16 + dir]; }
            if (slot) { if (*slot >= 0) { err = "duplicate state role"; return AUGB200_ERR_UNSUPPORTED; } *slot = (int8_t)s; }
            continue;
        }
        switch (sd.kind) {
        case K_IGENIC: sd.chain = 0; slot = &m.chain_state[0]; break;
        case K_GEO: sd.chain = (int8_t)(1 + dir * 3 + f); slot = &m.chain_state[sd.chain]; break;
        case K_LONGDSS: slot = &m.r_longdss[dir][f]; break;
        case K_LESSD: slot = &m.r_lessd[dir][f]; break;
        case K_EQUALD: slot = &m.r_equald[dir][f]; break;
        case K_LONGASS: slot = &m.r_longass[dir][f]; break;
        case K_UTR:
            slot = &m.r_utr[dir][sd.u5][sd.uk];
            if (sd.uk == U_INTRON) sd.chain = (int8_t)(CH_UTR + dir * 2 + (sd.u5 ? 0 : 1));
            break;
        default:
            switch (sd.ek) {
            case E_SINGLE: slot = &m.r_single; break; case E_INITIAL: slot = &m.r_initial[f]; break;
            case E_INTERNAL: slot = &m.r_internal[f]; break; case E_TERMINAL: slot = &m.r_terminal; break;
            case E_RSINGLE: slot = &m.r_rsingle; break; case E_RINITIAL: slot = &m.r_rinitial; break;
            case E_RINTERNAL: slot = &m.r_rinternal[f]; break; default: slot = &m.r_rterminal[f];
            }
        }
        if (*slot >= 0) { err = "duplicate state role"; return AUGB200_ERR_UNSUPPORTED; }
        *slot = (int8_t)s;
        if (sd.chain >= CH_UTR) m.chain_state[sd.chain] = (int8_t)s;
    }
    if (m.utr) {
        /* UTR exon states: slot order of Sweep::process_column and the table-driven geometry of UtrModel::viterbiForwardAndSampling
         * (utrmodel.cc:821-916), getEndPositions (:1572-1643) and notEndPartEmiProb (:1173-1404) */
        const int TUW = m.tuw, TE = m.tss_end, TIW = m.tiw, DW = m.dss_start + m.dss_end + 2, AW = m.ass_start + m.ass_end + 2, UP = m.ass_up,
                  DE = m.dss_end, AS = m.ass_start, AE = m.ass_end, PB = m.boxlen, DPC = m.dpc;
        static const int kinds[4] = {U_SINGLE, U_INIT, U_INTERNAL, U_TERM};
        for (int dir = 0; dir < 2; dir++) for (int three = 0; three < 2; three++) for (int q = 0; q < 4; q++) {
            const int slot = dir * 8 + three * 4 + q, uk = kinds[q], st = m.r_utr[dir][three ? 0 : 1][uk];
            UtrDesc& u = m.ud[slot]; memset(&u, 0, sizeof u);
            m.uslot[slot] = (int8_t)st;
            if (st < 0) continue;
            m.st[st].ek = (int8_t)slot;
            u.ldi = (int8_t)(three * 4 + q);
            const int term5rm = (-UP - AW + TIW + AE < 0) ? UP + AW - TIW - AE : UP + AW;
            const int single5rm = TUW + 1 - TIW > 1 ? TUW + 1 - TIW : 1;
            if (!dir && !three) {
                switch (uk) {
                case U_SINGLE: u = UtrDesc{CL_T5, BS_TSSF, UE_ATG, US_INIT5, 1, 1, (int16_t)(TUW + TE), (int16_t)TUW, (int16_t)(m.umax - TIW + TUW), (int16_t)single5rm, (int16_t)TIW, 1, 0, u.ldi}; break;
                case U_INIT: u = UtrDesc{CL_T5, BS_TSSF, UE_DSSF, US_INIT5, 0, 1, (int16_t)(TUW + TE), (int16_t)TUW, (int16_t)(m.umax + 2 + DE + TUW), (int16_t)(TUW + TE + DW), (int16_t)(-(DE + 2)), (int16_t)(-DW + 1), 0, u.ldi}; break;
                case U_INTERNAL: u = UtrDesc{CL_A5, BS_ASSF, UE_DSSF, US_5, 0, 0, (int16_t)(UP + AW), (int16_t)(UP + AS + 2), (int16_t)(m.umax + 2 + DE + UP + AS + 2), (int16_t)(DW + UP + AW), (int16_t)(-(DE + 2)), (int16_t)(-DW + 1), 0, u.ldi}; break;
                default: u = UtrDesc{CL_A5, BS_ASSF, UE_ATG, US_5, 2, 0, (int16_t)(UP + AW), (int16_t)(UP + AS + 2), (int16_t)(m.umax - TIW + UP + AS + 2), (int16_t)term5rm, (int16_t)TIW, 1, 0, u.ldi};
                }
            } else if (!dir) {
                switch (uk) {
                case U_SINGLE: u = UtrDesc{CL_X3, BS_NONE, UE_TTSF, US_3, 0, 0, 0, 0, (int16_t)m.umax3s, (int16_t)(DPC + PB), 0, (int16_t)(-DPC - PB + 1), 0, u.ldi}; break;
                case U_INIT: u = UtrDesc{CL_X3, BS_NONE, UE_DSSF, US_3, 2, 0, 0, 0, (int16_t)(m.umax + 2 + DE), (int16_t)(DE + 2), (int16_t)(-(DE + 2)), (int16_t)(-DW + 1), 0, u.ldi}; break;
                case U_INTERNAL: u = UtrDesc{CL_A3, BS_ASSF, UE_DSSF, US_3, 0, 0, (int16_t)(UP + AW), (int16_t)(UP + AS + 2), (int16_t)(m.umax + 2 + DE + 2 + AS + UP), (int16_t)(DW + UP + AW), (int16_t)(-(DE + 2)), (int16_t)(-DW + 1), 0, u.ldi}; break;
                default: u = UtrDesc{CL_A3, BS_ASSF, UE_TTSF, US_3, 0, 0, (int16_t)(UP + AW), (int16_t)(UP + AS + 2), (int16_t)(m.umax3t + 2 + AS + UP), (int16_t)(DPC + PB + AW + UP), 0, (int16_t)(-DPC - PB + 1), 0, u.ldi};
                }
            } else if (!three) {
                switch (uk) {
                case U_SINGLE: u = UtrDesc{CL_XRS, BS_NONE, UE_TSSR, US_RINIT5, 1, 0, 0, (int16_t)(-TIW), (int16_t)(m.umax - TIW + TUW), (int16_t)single5rm, (int16_t)(-TUW), (int16_t)(-TUW - TE + 1), 0, u.ldi}; break;
                case U_INIT: u = UtrDesc{CL_R5I, BS_DSSR, UE_TSSR, US_RINIT5, 0, 0, (int16_t)DW, (int16_t)(DE + 2), (int16_t)(m.umax + 2 + DE + TUW), (int16_t)(TUW + TE + DW), (int16_t)(-TUW), (int16_t)(-TUW - TE + 1), 0, u.ldi}; break;
                case U_INTERNAL: u = UtrDesc{CL_R5N, BS_DSSR, UE_ASSR, US_R5, 0, 0, (int16_t)DW, (int16_t)(DE + 2), (int16_t)(m.umax + 2 + DE + UP + AS + 2), (int16_t)(DW + UP + AW), (int16_t)(-(UP + AS + 2)), (int16_t)(-AW - UP + 1), 0, u.ldi}; break;
                default: u = UtrDesc{CL_XRT, BS_NONE, UE_ASSR, US_R5, 2, 0, 0, (int16_t)(-TIW), (int16_t)(m.umax - TIW + UP + AS + 2), (int16_t)term5rm, (int16_t)(-(UP + AS + 2)), (int16_t)(-AW - UP + 1), 0, u.ldi};
                }
            } else {
                switch (uk) {
                case U_SINGLE: u = UtrDesc{CL_TR, BS_TTSR, UE_RSTOP, US_R3, 0, 1, (int16_t)(PB + DPC), 0, (int16_t)m.umax3s, (int16_t)(DPC + PB), 0, 1, 0, u.ldi}; break;
                case U_INIT: u = UtrDesc{CL_R3, BS_DSSR, UE_RSTOP, US_R3, 2, 0, (int16_t)DW, (int16_t)(DE + 2), (int16_t)(m.umax + 2 + DE), (int16_t)(DE + 2), 0, 1, 0, u.ldi}; break;
                case U_INTERNAL: u = UtrDesc{CL_R3, BS_DSSR, UE_ASSR, US_R3, 0, 0, (int16_t)DW, (int16_t)(DE + 2), (int16_t)(m.umax + 2 + DE + UP + AS + 2), (int16_t)(DW + UP + AW), (int16_t)(-(UP + AS + 2)), (int16_t)(-AW - UP + 1), 0, u.ldi}; break;
                default: u = UtrDesc{CL_TR, BS_TTSR, UE_ASSR, US_R3, 0, 1, (int16_t)(PB + DPC), 0, (int16_t)(m.umax3t + 2 + AS + UP), (int16_t)(DPC + PB + AW + UP), (int16_t)(-(UP + AS + 2)), (int16_t)(-AW - UP + 1), 0, u.ldi};
                }
            }
            /* every endOfPred of the loop must index the length table: lm_off bounds the length */
            if (u.rm_off < 1) { err = "UTR geometry lets a state end before it begins"; return AUGB200_ERR_UNSUPPORTED; }
        }
        if (m.nc) {
            /* ncinternal: acceptor site after ncintron ... donor site; rncinternal: reverse donor after rncintron ... reverse acceptor.
             * NcModel::viterbiForwardAndSampling (ncmodel.cc:186-235): rightMost as for a UTR internal exon with minimum length 1, and
             * without exon hints leftMost = rightMost - 200; getEndPositions :702-725; notEndPartEmiProb :471-479, :505-513; the
             * content is NcModel::segProbs (intron emissions), the length distribution the coding internal one (:146) */
            const int16_t rm = (int16_t)(DE + UP + AS + 5);
            m.ud[16] = UtrDesc{CL_NCA, BS_ASSF, UE_DSSF, US_NC, 0, 0, (int16_t)(UP + AW), (int16_t)(UP + AS + 2), (int16_t)(rm + 200), rm, (int16_t)(-(DE + 2)), (int16_t)(-DW + 1), 0, 10};
            m.ud[17] = UtrDesc{CL_NCR, BS_DSSR, UE_ASSR, US_NC, 0, 0, (int16_t)DW, (int16_t)(DE + 2), (int16_t)(rm + 200), rm, (int16_t)(-(UP + AS + 2)), (int16_t)(-AW - UP + 1), 0, 10};
            bool haven = false;
            for (int ch = CH_NC; ch < NCHAIN; ch++) {
                int cs = m.chain_state[ch]; if (cs < 0) continue;
                for (int c = 0; c < m.C; c++) {
                    sc_t t = T[((size_t)c * m.S + cs) * m.S + cs];
                    if (!haven) { m.nc_tself = t; haven = true; } else if (t != m.nc_tself) { err = "nc intron self-loops differ"; return AUGB200_ERR_UNSUPPORTED; }
                }
            }
        }
        /* the four UTR-intron chains share the intron emission prefix and need one class-independent self transition */
        bool have = false;
        for (int ch = CH_UTR; ch < CH_NC; ch++) {
            int cs = m.chain_state[ch]; if (cs < 0) continue;
            for (int c = 0; c < m.C; c++) {
                sc_t t = T[((size_t)c * m.S + cs) * m.S + cs];
                if (!have) { m.utr_tself = t; have = true; } else if (t != m.utr_tself) { err = "UTR intron self-loops differ"; return AUGB200_ERR_UNSUPPORTED; }
            }
        }
    }
    {
        int q = 0;
        m.xslot[q++] = m.r_single; m.xslot[q++] = m.r_terminal;
        for (int f = 0; f < 3; f++) m.xslot[q++] = m.r_initial[f];
        for (int f = 0; f < 3; f++) m.xslot[q++] = m.r_internal[f];
        m.xslot[q++] = m.r_rsingle; m.xslot[q++] = m.r_rinitial;
        for (int f = 0; f < 3; f++) m.xslot[q++] = m.r_rinternal[f];
        for (int f = 0; f < 3; f++) m.xslot[q++] = m.r_rterminal[f];
    }
    if (m.chain_state[0] < 0) { err = "model has no intergenic state"; return AUGB200_ERR_UNSUPPORTED; }
    /* topology the kernels rely on (config/model/trans_shadow_*.pbl); anything else is rejected, not approximated */
    auto kind_of = [&](int a) { return m.st[a].kind; };
    for (int s = 0; s < m.S; s++) {
        StateDesc& sd = m.st[s];
        bool ok = true; bool selfloop = false;
        for (int i = 0; i < sd.nanc; i++) {
            int a = sd.anc[i]; const StateDesc& ad = m.st[a];
            if (a == s) { selfloop = true; continue; }
            switch (sd.kind) {
            case K_IGENIC: ok &= ad.kind == K_EXON || (ad.kind == K_UTR && ad.uk != U_INTRON) || ad.kind == K_DEAD; break;
            case K_DEAD: break;                    /* never evaluated: only its column-0 cell exists */
            case K_UTR: {
                if (sd.u5 == 2) {              /* nc: the chain is entered from nc exon states, the internal exon follows its chain (list of splice sites) */
                    if (sd.uk == U_INTRON) ok &= ad.u5 == 2 && ad.fwd == sd.fwd && (ad.kind == K_DEAD || ad.uk == U_INTERNAL);
                    else ok &= ad.u5 == 2 && ad.fwd == sd.fwd && (ad.kind == K_DEAD || ad.uk == U_INTRON);
                    break;
                }
                /* the predecessor kinds the candidate lists of Sweep::utr_eval are built for; intronvar states (hint-only) never hold a cell */
                if (ad.kind == K_UTR && ad.uk == U_INTRONVAR) break;
                if (sd.uk == U_INTRONVAR) break;
                if (sd.uk == U_INTRON) { ok &= ad.kind == K_UTR && ad.fwd == sd.fwd && ad.u5 == sd.u5 && ad.uk != U_INTRON; break; }
                const UtrDesc& u = m.ud[sd.ek];
                switch (u.list) {
                case CL_T5: case CL_TR: ok &= ad.kind == K_IGENIC; break;
                case CL_A5: case CL_A3: case CL_R5I: case CL_R5N: case CL_R3: ok &= ad.kind == K_UTR && ad.uk == U_INTRON && ad.fwd == sd.fwd && ad.u5 == sd.u5; break;
                case CL_X3: ok &= ad.kind == K_EXON && (ad.ek == E_SINGLE || ad.ek == E_TERMINAL); break;
                default: ok &= ad.kind == K_EXON && (ad.ek == E_RSINGLE || ad.ek == E_RINITIAL);
                }
                break;
            }
            case K_GEO: ok &= ad.kind == K_EQUALD && ad.fwd == sd.fwd && ad.frame == sd.frame; break;
            case K_EQUALD: case K_LESSD:
                ok &= (sd.fwd ? ad.kind == K_LONGDSS : ad.kind == K_LONGASS) && ad.fwd == sd.fwd && ad.frame == sd.frame && sd.nanc == 1; break;
            case K_LONGDSS: ok &= sd.fwd ? ad.kind == K_EXON : (ad.kind == K_LESSD || ad.kind == K_GEO); break;
            case K_LONGASS: ok &= sd.fwd ? (ad.kind == K_LESSD || ad.kind == K_GEO) : ad.kind == K_EXON; break;
            default:
                switch (sd.ek) {
                case E_INTERNAL: case E_TERMINAL: ok &= ad.kind == K_LONGASS && ad.fwd; break;
                case E_RINTERNAL: case E_RINITIAL: ok &= ad.kind == K_LONGDSS && !ad.fwd; break;
                case E_SINGLE: case E_INITIAL:      /* igenic, or (UTR models) the 5' UTR states that end before the start codon */
                    ok &= m.utr ? (ad.kind == K_UTR && ad.fwd && ad.u5 && (ad.uk == U_SINGLE || ad.uk == U_TERM)) : (ad.kind == K_IGENIC && sd.nanc == 1); break;
                default:                            /* rsingle, rterminal */
                    ok &= m.utr ? (ad.kind == K_UTR && !ad.fwd && !ad.u5 && (ad.uk == U_SINGLE || ad.uk == U_INIT)) : (ad.kind == K_IGENIC && sd.nanc == 1);
                }
            }
        }
        if (sd.kind == K_UTR && sd.uk == U_INTRONVAR) { if (!isneg(tab[o_init + s])) ok = false; }     /* must stay empty */
        if (sd.kind == K_DEAD && sd.uk == U_INTRONVAR) { if (!isneg(tab[o_init + s])) ok = false; }
        /* a dead state with an initial probability may only lead into chains (column 0 -> column 1) or other dead states: a table-driven
         * exon state reads its predecessors from a site list, which holds chain values only */
        if (sd.kind == K_DEAD)
            for (int b2 = 0; b2 < m.S; b2++)
                if (!isneg(T[(size_t)s * m.S + b2]) && m.st[b2].kind != K_DEAD && m.st[b2].chain < 0 && !isneg(tab[o_init + s])) {
                    /* e.g. ncintronvar -> ncinternal: fine as long as the dead state itself starts empty */
                    ok = false;
                }
        if (selfloop != (sd.chain >= 0)) ok = false;
        if (!ok) { err = "unsupported transition topology at state " + std::to_string(s); return AUGB200_ERR_UNSUPPORTED; }
        (void)kind_of;
        /* one-base successors of a cell: which chain does this state enter? */
        for (int c = 0; c < NCHAIN; c++) {
            int cs = m.chain_state[c]; if (cs < 0 || cs == s) continue;
            if (!isneg(T[(size_t)s * m.S + cs])) { if (sd.feeds >= 0) { err = "state feeds two chains"; return AUGB200_ERR_UNSUPPORTED; } sd.feeds = (int8_t)c; }
        }
    }
    /* the 6 geometric chains share one prefix array: emission and self-loop must agree per class */
    for (int c = 0; c < m.C; c++) {
        sc_t ref = 0; bool have = false;
        for (int ch = 1; ch < CH_UTR; ch++) {
            int cs = m.chain_state[ch]; if (cs < 0) continue;
            sc_t t = T[((size_t)c * m.S + cs) * m.S + cs];
            if (!have) { ref = t; have = true; } else if (t != ref) { err = "geometric self-loops differ"; return AUGB200_ERR_UNSUPPORTED; }
        }
    }
    /* ---- pointers ---- */
    const sc_t* b = tab.data();
    m.init = b + o_init; m.term = b + o_term; m.trans = b + o_trans;
    m.xemi = b + o_xemi; m.xinit = b + o_xinit; m.xet = b + o_xet;
    for (int l = 0; l < 5; l++) m.xpls[l] = l <= m.k ? b + o_xpls[l] : nullptr;
    m.iemi = b + o_iemi; m.gemi = b + o_gemi; m.gfirst = b + o_gfirst; m.tis = b + o_tis; m.assm = b + o_assm;
    m.tis_bb = m.tis_nbins > 1 ? b + o_tbb : nullptr; m.tis_bp = m.tis_nbins > 0 ? b + o_tbp : nullptr;
    m.ld_single = b + o_lds; m.ld_initial = b + o_ldi; m.ld_internal = b + o_ldn; m.ld_terminal = b + o_ldt; m.ld_intron = b + o_ldx;
    m.ass_pat = b + o_ap; m.ass_pat_non = b + o_apn; m.dss_pat = b + o_dp; m.dss_pat_non = b + o_dpn;
    if (m.utr) {
        m.u5i = b + o_u5i; m.u5 = b + o_u5; m.u3 = b + o_u3; m.tup = b + o_tup;
        m.tssm = b + o_tssm; m.tsstm = b + o_tsstm; m.tatam = b + o_tatam; m.ttsm = b + o_ttsm; m.aataaa = b + o_aat;
        for (int i = 0; i < 10; i++) m.uld[i] = b + o_uld[i];
        m.uld[10] = m.ld_internal; m.n_uld[10] = m.n_ld_exon;          /* NcModel::lenDistInternal = ExonModel::lenDistInternal (ncmodel.cc:146) */
    }
    return AUGB200_OK;
}

DevModel HostModel::rebased(const sc_t* base) const {
    DevModel r = dm; const sc_t* b = tab.data();
    auto rb = [&](const sc_t* p) { return p ? base + (p - b) : nullptr; };
    r.init = rb(dm.init); r.term = rb(dm.term); r.trans = rb(dm.trans); r.xemi = rb(dm.xemi); r.xinit = rb(dm.xinit); r.xet = rb(dm.xet);
    for (int l = 0; l < 5; l++) r.xpls[l] = rb(dm.xpls[l]);
    r.iemi = rb(dm.iemi); r.gemi = rb(dm.gemi); r.gfirst = rb(dm.gfirst); r.tis = rb(dm.tis); r.assm = rb(dm.assm);
    r.tis_bb = rb(dm.tis_bb); r.tis_bp = rb(dm.tis_bp);
    r.ld_single = rb(dm.ld_single); r.ld_initial = rb(dm.ld_initial); r.ld_internal = rb(dm.ld_internal); r.ld_terminal = rb(dm.ld_terminal); r.ld_intron = rb(dm.ld_intron);
    r.ass_pat = rb(dm.ass_pat); r.ass_pat_non = rb(dm.ass_pat_non); r.dss_pat = rb(dm.dss_pat); r.dss_pat_non = rb(dm.dss_pat_non);
    r.u5i = rb(dm.u5i); r.u5 = rb(dm.u5); r.u3 = rb(dm.u3); r.tup = rb(dm.tup);
    r.tssm = rb(dm.tssm); r.tsstm = rb(dm.tsstm); r.tatam = rb(dm.tatam); r.ttsm = rb(dm.ttsm); r.aataaa = rb(dm.aataaa);
    for (int i = 0; i < 11; i++) r.uld[i] = rb(dm.uld[i]);
    return r;
}

}  // namespace augb
