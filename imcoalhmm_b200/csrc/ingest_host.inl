// Ingest of alignments (included by imc_lib.cu): the pairwise symbol rule, a FASTA reader and the binary container.

// ------------------------------------------------------------------------------------------ ingest
// Pairwise symbol rule of the reference's preprocessing script (scripts/prepare-alignments.py:99-111):
//   upper-case both bases; 2 if either is not one of A, C, G, T; 0 if equal; 1 otherwise.
static inline uint8_t pair_symbol(unsigned char a, unsigned char b) {
    static const struct Table { uint8_t code[256]; Table() {
        for (int i = 0; i < 256; ++i) code[i] = 4;
        code[(int)'A'] = code[(int)'a'] = 0; code[(int)'C'] = code[(int)'c'] = 1;
        code[(int)'G'] = code[(int)'g'] = 2; code[(int)'T'] = code[(int)'t'] = 3;
    } } tab;
    const uint8_t x = tab.code[a], y = tab.code[b];
    return (x > 3 || y > 3) ? 2 : (x == y ? 0 : 1);
}

extern "C" int imc_seq_from_pair(const char* seq1, const char* seq2, int64_t L, imc_seq** out) {
    if (!out || L < 0 || (L > 0 && (!seq1 || !seq2))) return fail(IMC_ERR_INVALID, "bad arguments");
    imc_seq* s = new (std::nothrow) imc_seq;
    if (!s) return fail(IMC_ERR_NOMEM, "out of memory");
    s->nsym = 3;
    try { s->sym.resize((size_t)L); } catch (...) { delete s; return fail(IMC_ERR_NOMEM, "out of memory for %lld symbols", (long long)L); }
    for (int64_t t = 0; t < L; ++t) s->sym[(size_t)t] = pair_symbol((unsigned char)seq1[t], (unsigned char)seq2[t]);
    return seq_finish(s, out);
}

// Triplet / quartet columns (scripts/prepare-alignments.py:113-190): with every base in ACGT the symbol is
// i1 + 4 i2 + 16 i3 (+ 32 i4 for quartets -- the script's own weight, reproduced as it is, so clean quartet columns
// reach 159 and 128 is ambiguous: NSYM = 160 there), otherwise 64 (128).
extern "C" int imc_seq_from_columns(const char* const* seqs, int n_seqs, int64_t L, imc_seq** out) {
    if (!out || !seqs || L < 0) return fail(IMC_ERR_INVALID, "bad arguments");
    if (n_seqs == 2) return imc_seq_from_pair(seqs[0], seqs[1], L, out);
    if (n_seqs != 3 && n_seqs != 4) return fail(IMC_ERR_INVALID, "alignments of 2, 3 or 4 sequences are supported, not %d", n_seqs);
    for (int k = 0; k < n_seqs; ++k) if (L > 0 && !seqs[k]) return fail(IMC_ERR_INVALID, "sequence %d is NULL", k);
    static const struct Table { uint8_t code[256]; Table() {
        for (int i = 0; i < 256; ++i) code[i] = 4;
        code[(int)'A'] = code[(int)'a'] = 0; code[(int)'C'] = code[(int)'c'] = 1;
        code[(int)'G'] = code[(int)'g'] = 2; code[(int)'T'] = code[(int)'t'] = 3;
    } } tab;
    static const int weight[4] = {1, 4, 16, 32};
    const int missing = n_seqs == 3 ? 64 : 128;
    imc_seq* s = new (std::nothrow) imc_seq;
    if (!s) return fail(IMC_ERR_NOMEM, "out of memory");
    s->nsym = n_seqs == 3 ? 65 : 160;
    try { s->sym.resize((size_t)L); } catch (...) { delete s; return fail(IMC_ERR_NOMEM, "out of memory for %lld symbols", (long long)L); }
    for (int64_t t = 0; t < L; ++t) {
        int v = 0;
        bool clean = true;
        for (int k = 0; k < n_seqs; ++k) {
            const uint8_t c = tab.code[(unsigned char)seqs[k][t]];
            clean = clean && c < 4;
            v += weight[k] * (c & 3);
        }
        s->sym[(size_t)t] = (uint8_t)(clean ? v : missing);
    }
    return seq_finish(s, out);
}

// FASTA: '>' starts a record, its name is the text up to the first whitespace; sequence lines are concatenated.
static int read_fasta(const char* path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IMC_ERR_IO, "cannot open '%s': %s", path, strerror(errno));
    std::vector<char> buf(1 << 20);
    bool in_header = false, at_line_start = true;
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got; ++i) {
            const char ch = buf[i];
            if (in_header) {
                if (ch == '\n') { in_header = false; at_line_start = true; }
                else names.back().push_back(ch);
                continue;
            }
            if (ch == '\n' || ch == '\r') { at_line_start = (ch == '\n') || at_line_start; continue; }
            if (at_line_start && ch == '>') { names.emplace_back(); seqs.emplace_back(); in_header = true; continue; }
            at_line_start = false;
            if (ch == ' ' || ch == '\t') continue;
            if (seqs.empty()) { fclose(f); return fail(IMC_ERR_IO, "'%s' does not start with a FASTA header", path); }
            seqs.back().push_back(ch);
        }
    }
    fclose(f);
    for (auto& n : names) {
        while (!n.empty() && (n.back() == '\r' || n.back() == ' ' || n.back() == '\t')) n.pop_back();
        const size_t sp = n.find_first_of(" \t");
        if (sp != std::string::npos) n.resize(sp);
    }
    return IMC_OK;
}

extern "C" int imc_seq_from_fasta(const char* path, const char* name1, const char* name2, imc_seq** out) {
    if (!path || !out) return fail(IMC_ERR_INVALID, "NULL argument");
    std::vector<std::string> names, seqs;
    int rc = read_fasta(path, names, seqs);
    if (rc) return rc;
    int i1 = -1, i2 = -1;
    if (!name1 && !name2) {
        if (names.size() != 2) return fail(IMC_ERR_INVALID, "'%s' holds %zu records; name the two to compare", path, names.size());
        i1 = 0; i2 = 1;
    } else {
        if (!name1 || !name2) return fail(IMC_ERR_INVALID, "give both record names or neither");
        for (size_t i = 0; i < names.size(); ++i) { if (names[i] == name1) i1 = (int)i; if (names[i] == name2) i2 = (int)i; }
        if (i1 < 0 || i2 < 0) return fail(IMC_ERR_INVALID, "record '%s' not found in '%s'", i1 < 0 ? name1 : name2, path);
    }
    if (seqs[i1].size() != seqs[i2].size())
        return fail(IMC_ERR_INVALID, "aligned sequences differ in length (%zu vs %zu)", seqs[i1].size(), seqs[i2].size());
    return imc_seq_from_pair(seqs[i1].data(), seqs[i2].data(), (int64_t)seqs[i1].size(), out);
}

extern "C" int imc_seq_from_fasta_n(const char* path, const char* const* record_names, int n_names, imc_seq** out) {
    if (!path || !out || !record_names) return fail(IMC_ERR_INVALID, "NULL argument");
    if (n_names < 2 || n_names > 4) return fail(IMC_ERR_INVALID, "2, 3 or 4 record names are needed, got %d", n_names);
    std::vector<std::string> names, seqs;
    int rc = read_fasta(path, names, seqs);
    if (rc) return rc;
    const char* cols[4];
    size_t len = 0;
    for (int k = 0; k < n_names; ++k) {
        int found = -1;
        for (size_t i = 0; i < names.size(); ++i) if (record_names[k] && names[i] == record_names[k]) found = (int)i;
        if (found < 0) return fail(IMC_ERR_INVALID, "record '%s' not found in '%s'", record_names[k] ? record_names[k] : "(null)", path);
        if (k > 0 && seqs[found].size() != len)
            return fail(IMC_ERR_INVALID, "aligned sequences differ in length (%zu vs %zu)", len, seqs[found].size());
        len = seqs[found].size();
        cols[k] = seqs[found].data();
    }
    return imc_seq_from_columns(cols, n_names, (int64_t)len, out);
}

// Binary container: "IMCSEQ1\0", int32 nsym, int32 bits per symbol (2 or 8), int64 L, packed symbols (little endian,
// symbol t of a 2-bit file sits in bits 2*(t%4).. of byte t/4).  16x smaller than the text format for NSYM = 3.
extern "C" int imc_seq_save(const imc_seq* seq, const char* path) {
    if (!seq || !path) return fail(IMC_ERR_INVALID, "NULL argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(IMC_ERR_IO, "cannot create '%s': %s", path, strerror(errno));
    const int32_t nsym = seq->nsym, bits = seq->nsym <= 4 ? 2 : 8;
    const int64_t L = (int64_t)seq->sym.size();
    bool ok = fwrite("IMCSEQ1", 1, 8, f) == 8 && fwrite(&nsym, 4, 1, f) == 1 && fwrite(&bits, 4, 1, f) == 1 && fwrite(&L, 8, 1, f) == 1;
    if (ok && bits == 8) ok = L == 0 || fwrite(seq->sym.data(), 1, (size_t)L, f) == (size_t)L;
    if (ok && bits == 2) {
        std::vector<uint8_t> packed((size_t)((L + 3) / 4), 0);
        for (int64_t t = 0; t < L; ++t) packed[(size_t)(t >> 2)] |= (uint8_t)(seq->sym[(size_t)t] << (2 * (t & 3)));
        ok = packed.empty() || fwrite(packed.data(), 1, packed.size(), f) == packed.size();
    }
    if (fclose(f) != 0) ok = false;
    return ok ? IMC_OK : fail(IMC_ERR_IO, "short write to '%s'", path);
}

extern "C" int imc_seq_load(const char* path, imc_seq** out) {
    if (!path || !out) return fail(IMC_ERR_INVALID, "NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IMC_ERR_IO, "cannot open '%s': %s", path, strerror(errno));
    char magic[8];
    int32_t nsym = 0, bits = 0;
    int64_t L = -1;
    bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, "IMCSEQ1", 8) && fread(&nsym, 4, 1, f) == 1 && fread(&bits, 4, 1, f) == 1 &&
              fread(&L, 8, 1, f) == 1 && nsym >= 1 && nsym <= 255 && (bits == 2 || bits == 8) && L >= 0 && !(bits == 2 && nsym > 4);
    if (!ok) { fclose(f); return fail(IMC_ERR_IO, "'%s' is not an IMCSEQ1 file", path); }
    imc_seq* s = new (std::nothrow) imc_seq;
    if (!s) { fclose(f); return fail(IMC_ERR_NOMEM, "out of memory"); }
    s->nsym = nsym;
    try {
        s->sym.resize((size_t)L);
        if (bits == 8) ok = L == 0 || fread(s->sym.data(), 1, (size_t)L, f) == (size_t)L;
        else {
            std::vector<uint8_t> packed((size_t)((L + 3) / 4));
            ok = packed.empty() || fread(packed.data(), 1, packed.size(), f) == packed.size();
            for (int64_t t = 0; ok && t < L; ++t) s->sym[(size_t)t] = (packed[(size_t)(t >> 2)] >> (2 * (t & 3))) & 3;
        }
    } catch (...) { fclose(f); delete s; return fail(IMC_ERR_NOMEM, "out of memory for %lld symbols", (long long)L); }
    fclose(f);
    if (ok) for (uint8_t v : s->sym) if (v >= nsym) { ok = false; break; }
    if (!ok) { delete s; return fail(IMC_ERR_IO, "'%s' is truncated or holds symbols outside [0, %d)", path, nsym); }
    return seq_finish(s, out);
}

