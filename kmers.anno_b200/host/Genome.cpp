// Genome.cpp — GTO (JSON) / FASTA readers for the `apply` path.  See Genome.hpp.
#include "Genome.hpp"

#include <algorithm>
#include <dirent.h>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace theseed {

std::string readFile(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw FileNotFoundException("File " + path + " not found or unreadable.");
    std::ostringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

std::vector<std::string> readLines(const std::string& path) {
    std::string text = readFile(path);
    std::vector<std::string> out;
    size_t i = 0;
    while (i < text.size()) {
        size_t e = text.find('\n', i);
        if (e == std::string::npos) e = text.size();
        size_t len = e - i;
        if (len && text[i + len - 1] == '\r') len--;
        out.emplace_back(text, i, len);
        i = e + 1;
    }
    return out;
}

namespace {

// Minimal JSON reader: walks the document, materialising only the strings asked for.
class Json {
public:
    explicit Json(const std::string& t) : s(t) {}
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++; }
    char peek() { ws(); if (p >= s.size()) fail("unexpected end"); return s[p]; }
    void expect(char c) { if (peek() != c) fail(std::string("expected '") + c + "'"); p++; }
    bool consume(char c) { if (peek() == c) { p++; return true; } return false; }
    [[noreturn]] void fail(const std::string& m) { throw IOException("JSON error at byte " + std::to_string(p) + ": " + m); }

    std::string str() {
        expect('"');
        std::string out;
        while (p < s.size() && s[p] != '"') {
            char c = s[p++];
            if (c != '\\') { out.push_back(c); continue; }
            if (p >= s.size()) fail("bad escape");
            char e = s[p++];
            switch (e) {
                case 'n': out.push_back('\n'); break;
                case 't': out.push_back('\t'); break;
                case 'r': out.push_back('\r'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'u': {
                    if (p + 4 > s.size()) fail("bad \\u escape");
                    unsigned cp = (unsigned)std::stoul(s.substr(p, 4), nullptr, 16);
                    p += 4;
                    if (cp < 0x80) out.push_back((char)cp);
                    else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                    else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                    break;
                }
                default: out.push_back(e);
            }
        }
        if (p >= s.size()) fail("unterminated string");
        p++;
        return out;
    }
    void skipString() {
        expect('"');
        while (p < s.size() && s[p] != '"') p += (s[p] == '\\') ? 2 : 1;
        if (p >= s.size()) fail("unterminated string");
        p++;
    }
    void skipValue() {
        char c = peek();
        if (c == '"') skipString();
        else if (c == '{') { p++; if (consume('}')) return; do { skipString(); expect(':'); skipValue(); } while (consume(',')); expect('}'); }
        else if (c == '[') { p++; if (consume(']')) return; do { skipValue(); } while (consume(',')); expect(']'); }
        else { while (p < s.size() && s[p] != ',' && s[p] != '}' && s[p] != ']' && s[p] != ' ' && s[p] != '\n' && s[p] != '\r' && s[p] != '\t') p++; }
    }
    /** string value, or "" for null / non-string values (which are skipped) */
    std::string strOrEmpty() { if (peek() == '"') return str(); skipValue(); return ""; }

    const std::string& s;
    size_t p = 0;
};

bool endsWith(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

}  // namespace

Genome::Genome(const std::string& path) {
    std::string text = readFile(path);
    size_t slash = path.find_last_of('/');
    std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
    if (endsWith(base, ".gto")) loadGto(text);
    else { fromFasta_ = true; loadFasta(text, base.substr(0, base.find_last_of('.'))); }
}

void Genome::loadGto(const std::string& text) {
    Json j(text);
    j.expect('{');
    if (j.consume('}')) return;
    do {
        std::string key = j.str();
        j.expect(':');
        if (key == "id") id_ = j.strOrEmpty();
        else if (key == "scientific_name") name_ = j.strOrEmpty();
        else if (key == "features") {
            j.expect('[');
            if (!j.consume(']')) {
                do {
                    std::string fid, type, fun, prot;
                    j.expect('{');
                    if (!j.consume('}')) {
                        do {
                            std::string k = j.str();
                            j.expect(':');
                            if (k == "id") fid = j.strOrEmpty();
                            else if (k == "type") type = j.strOrEmpty();
                            else if (k == "function") fun = j.strOrEmpty();
                            else if (k == "protein_translation") prot = j.strOrEmpty();
                            else j.skipValue();
                        } while (j.consume(','));
                        j.expect('}');
                    }
                    features_.emplace_back(std::move(fid), std::move(type), std::move(fun), std::move(prot));
                } while (j.consume(','));
                j.expect(']');
            }
        } else j.skipValue();
    } while (j.consume(','));
}

void Genome::loadFasta(const std::string& text, const std::string& stem) {
    id_ = stem;
    name_ = stem;
    size_t i = 0;
    std::string fid, fun, seq;
    bool have = false;
    auto flush = [&] {
        if (have) features_.emplace_back(fid, "CDS", fun, seq);
        have = false; seq.clear();
    };
    while (i < text.size()) {
        size_t e = text.find('\n', i);
        if (e == std::string::npos) e = text.size();
        size_t len = e - i;
        if (len && text[i + len - 1] == '\r') len--;
        if (len && text[i] == '>') {
            flush();
            std::string hdr(text, i + 1, len - 1);
            size_t sp = hdr.find_first_of(" \t");
            fid = hdr.substr(0, sp);
            fun = sp == std::string::npos ? "" : hdr.substr(hdr.find_first_not_of(" \t", sp));
            have = true;
        } else if (have) {
            seq.append(text, i, len);
        }
        i = e + 1;
    }
    flush();
}

std::vector<const Feature*> Genome::getPegs() const {
    std::vector<const Feature*> out;
    bool any_peg_id = false;
    for (const Feature& f : features_) if (f.isPeg()) { any_peg_id = true; break; }
    for (const Feature& f : features_) {
        // features are typed by their fid (Genome.getPegs()); only a protein FASTA without fig-style ids is all proteins
        if (f.isPeg() || (fromFasta_ && !any_peg_id)) out.push_back(&f);
    }
    return out;
}

GenomeDirectory::GenomeDirectory(const std::string& dir) {
    struct stat st;
    if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode))
        throw FileNotFoundException("Input directory " + dir + " not found or invalid.");
    DIR* d = opendir(dir.c_str());
    if (!d) throw FileNotFoundException("Input directory " + dir + " not found or invalid.");
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (endsWith(n, ".gto") || endsWith(n, ".faa") || endsWith(n, ".fa") || endsWith(n, ".fasta"))
            files_.push_back(dir + "/" + n);
    }
    closedir(d);
    std::sort(files_.begin(), files_.end());
}

}  // namespace theseed
