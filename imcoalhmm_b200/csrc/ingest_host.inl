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
struct FileCloser {          // closes the file on every exit path, exceptions included
    FILE* f;
    ~FileCloser() { if (f) fclose(f); }
};

// No C++ exception may cross the C ABI: the readers below run inside these wrappers (a header such as "999999999999 10" or
// a file larger than memory must come back as an error code, not terminate the caller's process).
static int read_fasta_impl(const char* path, std::vector<std::string>& names, std::vector<std::string>& seqs);
static int read_phylip_impl(const char* path, bool sequential, bool relaxed, std::vector<std::string>& names, std::vector<std::string>& seqs);
static int read_fasta(const char* path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    try { return read_fasta_impl(path, names, seqs); }
    catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of memory reading '%s'", path); }
    catch (const std::exception& e) { return fail(IMC_ERR_IO, "reading '%s' failed: %s", path, e.what()); }
    catch (...) { return fail(IMC_ERR_IO, "reading '%s' failed", path); }
}
static int read_phylip(const char* path, bool sequential, bool relaxed, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    try { return read_phylip_impl(path, sequential, relaxed, names, seqs); }
    catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of memory reading '%s'", path); }
    catch (const std::exception& e) { return fail(IMC_ERR_IO, "reading '%s' failed: %s", path, e.what()); }
    catch (...) { return fail(IMC_ERR_IO, "reading '%s' failed", path); }
}

static int read_fasta_impl(const char* path, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IMC_ERR_IO, "cannot open '%s': %s", path, strerror(errno));
    FileCloser closer{f};
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
            if (seqs.empty()) return fail(IMC_ERR_IO, "'%s' does not start with a FASTA header", path);
            seqs.back().push_back(ch);
        }
    }
    for (auto& n : names) {
        while (!n.empty() && (n.back() == '\r' || n.back() == ' ' || n.back() == '\t')) n.pop_back();
        const size_t sp = n.find_first_of(" \t");
        if (sp != std::string::npos) n.resize(sp);
    }
    return IMC_OK;
}

// PHYLIP (the other common input of scripts/prepare-alignments.py, whose <input format> argument is handed to BioPython):
// header "ntaxa nsites", then either interleaved blocks (first block: name + data per taxon; later blocks: data only, same
// order) or, sequential, each taxon's name followed by all of its data over as many lines as it takes.  Strict names
// occupy the first 10 columns of the line; relaxed names end at the first blank.  Blanks inside the data are ignored;
// '.' (match-first-row shorthand) is rejected, as BioPython does.
static int read_phylip_impl(const char* path, bool sequential, bool relaxed, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(IMC_ERR_IO, "cannot open '%s': %s", path, strerror(errno));
    std::string text;
    std::vector<char> buf(1 << 20);
    size_t got;
    try {
        while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) text.append(buf.data(), got);
    } catch (...) { fclose(f); return fail(IMC_ERR_NOMEM, "out of memory reading '%s'", path); }
    fclose(f);
    std::vector<std::pair<size_t, size_t>> lines;      // non-blank lines as (begin, end)
    for (size_t b = 0; b < text.size();) {
        size_t e = text.find('\n', b);
        if (e == std::string::npos) e = text.size();
        size_t e2 = e;
        while (e2 > b && (text[e2 - 1] == '\r' || text[e2 - 1] == ' ' || text[e2 - 1] == '\t')) --e2;
        bool blank = true;
        for (size_t i = b; i < e2; ++i) if (text[i] != ' ' && text[i] != '\t') { blank = false; break; }
        if (!blank) lines.emplace_back(b, e2);
        b = e + 1;
    }
    if (lines.empty()) return fail(IMC_ERR_IO, "'%s' is empty", path);
    long long ntaxa = 0, nsites = 0;
    {
        const std::string head = text.substr(lines[0].first, lines[0].second - lines[0].first);
        char extra;
        if (sscanf(head.c_str(), " %lld %lld %c", &ntaxa, &nsites, &extra) != 2 || ntaxa < 1 || nsites < 0)
            return fail(IMC_ERR_IO, "'%s' does not start with a PHYLIP header (taxa, sites)", path);
        // every taxon needs at least one line and every site at least one byte of the file
        if (ntaxa > (long long)lines.size() - 1 || nsites > (long long)text.size())
            return fail(IMC_ERR_IO, "'%s': header announces %lld taxa x %lld sites, more than the file can hold", path, ntaxa, nsites);
    }
    names.assign((size_t)ntaxa, std::string());
    seqs.assign((size_t)ntaxa, std::string());
    auto add_data = [&](std::string& dst, size_t b, size_t e) -> int {
        for (size_t i = b; i < e; ++i) {
            const char ch = text[i];
            if (ch == ' ' || ch == '\t') continue;
            if (ch == '.') return fail(IMC_ERR_IO, "'%s' uses '.' match characters, which are not supported", path);
            dst.push_back(ch);
        }
        return IMC_OK;
    };
    auto split_name = [&](size_t b, size_t e, std::string& name) -> size_t {      // returns where the data starts
        size_t i = b;
        if (relaxed) {
            while (i < e && (text[i] == ' ' || text[i] == '\t')) ++i;
            const size_t nb = i;
            while (i < e && text[i] != ' ' && text[i] != '\t') ++i;
            name = text.substr(nb, i - nb);
            return i;
        }
        const size_t ne = std::min(e, b + 10);
        name = text.substr(b, ne - b);
        while (!name.empty() && (name.back() == ' ' || name.back() == '\t')) name.pop_back();
        size_t lead = 0;
        while (lead < name.size() && (name[lead] == ' ' || name[lead] == '\t')) ++lead;
        name.erase(0, lead);
        return ne;
    };
    size_t li = 1;
    int rc;
    try {
        if (sequential) {
            for (long long t = 0; t < ntaxa; ++t) {
                if (li >= lines.size()) return fail(IMC_ERR_IO, "'%s' ends before taxon %lld", path, t + 1);
                size_t ds = split_name(lines[li].first, lines[li].second, names[(size_t)t]);
                if ((rc = add_data(seqs[(size_t)t], ds, lines[li].second))) return rc;
                ++li;
                while ((long long)seqs[(size_t)t].size() < nsites && li < lines.size()) {
                    if ((rc = add_data(seqs[(size_t)t], lines[li].first, lines[li].second))) return rc;
                    ++li;
                }
            }
        } else {
            for (long long t = 0; t < ntaxa; ++t, ++li) {
                if (li >= lines.size()) return fail(IMC_ERR_IO, "'%s' ends before taxon %lld", path, t + 1);
                size_t ds = split_name(lines[li].first, lines[li].second, names[(size_t)t]);
                if ((rc = add_data(seqs[(size_t)t], ds, lines[li].second))) return rc;
            }
            while (li < lines.size()) {
                for (long long t = 0; t < ntaxa; ++t, ++li) {
                    if (li >= lines.size()) return fail(IMC_ERR_IO, "'%s': the last interleaved block is incomplete", path);
                    if ((rc = add_data(seqs[(size_t)t], lines[li].first, lines[li].second))) return rc;
                }
            }
        }
    } catch (const std::bad_alloc&) { return fail(IMC_ERR_NOMEM, "out of memory reading '%s'", path); }
    if (li < lines.size()) return fail(IMC_ERR_IO, "'%s' holds data beyond the %lld x %lld announced in its header", path, ntaxa, nsites);
    for (long long t = 0; t < ntaxa; ++t)
        if ((long long)seqs[(size_t)t].size() != nsites)
            return fail(IMC_ERR_IO, "taxon '%s' of '%s' has %zu sites, the header says %lld", names[(size_t)t].c_str(), path,
                        seqs[(size_t)t].size(), nsites);
    return IMC_OK;
}

// format: "fasta", "phylip" (interleaved, 10-column names), "phylip-relaxed" (interleaved, names end at the first blank),
// "phylip-sequential" -- BioPython's names for them (prepare-alignments.py:41,66)
static int read_alignment(const char* path, const char* format, std::vector<std::string>& names, std::vector<std::string>& seqs) {
    if (!format || !strcmp(format, "fasta")) return read_fasta(path, names, seqs);
    if (!strcmp(format, "phylip")) return read_phylip(path, false, false, names, seqs);
    if (!strcmp(format, "phylip-relaxed")) return read_phylip(path, false, true, names, seqs);
    if (!strcmp(format, "phylip-sequential")) return read_phylip(path, true, false, names, seqs);
    return fail(IMC_ERR_UNSUPPORTED, "alignment format '%s' is not supported (fasta, phylip, phylip-relaxed, phylip-sequential)", format);
}

extern "C" int imc_seq_from_alignment(const char* path, const char* format, const char* const* record_names, int n_names, imc_seq** out) {
    if (!path || !out) return fail(IMC_ERR_INVALID, "NULL argument");
    if (n_names != 0 && (n_names < 2 || n_names > 4 || !record_names)) return fail(IMC_ERR_INVALID, "0 (the only two records), 2, 3 or 4 record names are needed");
    std::vector<std::string> names, seqs;
    int rc = read_alignment(path, format, names, seqs);
    if (rc) return rc;
    const char* cols[4];
    size_t len = 0;
    if (n_names == 0) {
        if (names.size() != 2) return fail(IMC_ERR_INVALID, "'%s' holds %zu records; name the ones to compare", path, names.size());
        cols[0] = seqs[0].data(); cols[1] = seqs[1].data(); len = seqs[0].size(); n_names = 2;
        if (seqs[1].size() != len) return fail(IMC_ERR_INVALID, "aligned sequences differ in length (%zu vs %zu)", len, seqs[1].size());
    } else {
        for (int k = 0; k < n_names; ++k) {
            int found = -1;
            for (size_t i = 0; i < names.size(); ++i) if (record_names[k] && names[i] == record_names[k]) found = (int)i;
            if (found < 0) return fail(IMC_ERR_INVALID, "record '%s' not found in '%s'", record_names[k] ? record_names[k] : "(null)", path);
            if (k > 0 && seqs[found].size() != len)
                return fail(IMC_ERR_INVALID, "aligned sequences differ in length (%zu vs %zu)", len, seqs[found].size());
            len = seqs[found].size();
            cols[k] = seqs[found].data();
        }
    }
    if (n_names == 2) return imc_seq_from_pair(cols[0], cols[1], (int64_t)len, out);
    return imc_seq_from_columns(cols, n_names, (int64_t)len, out);
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

