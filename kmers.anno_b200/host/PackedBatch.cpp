// PackedBatch.cpp — see PackedBatch.hpp.
#include "PackedBatch.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

namespace theseed {

namespace {

bool endsWith(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// FASTA in place: header = '>' id [blank function]; the sequence lines of a record are appended to
// `g.residues` without their terminators.  Same record rules as Genome::loadFasta.
void scanFasta(PackedGenome& g) {
    const char* t = g.text.data();
    const size_t n = g.text.size();
    g.residues.reserve(n);
    size_t i = 0;
    bool have = false;
    uint32_t len = 0;
    auto flush = [&] {
        if (have) g.lengths.push_back(len);
        have = false; len = 0;
    };
    while (i < n) {
        const char* nl = (const char*)memchr(t + i, '\n', n - i);
        const size_t e = nl ? (size_t)(nl - t) : n;
        size_t l = e - i;
        if (l && t[i + l - 1] == '\r') l--;
        if (l && t[i] == '>') {
            flush();
            size_t idEnd = i + 1;
            while (idEnd < i + l && t[idEnd] != ' ' && t[idEnd] != '\t') idEnd++;
            size_t fn = idEnd;
            while (fn < i + l && (t[fn] == ' ' || t[fn] == '\t')) fn++;
            g.pegs.push_back(PackedPeg{(uint32_t)(i + 1), (uint32_t)(idEnd - i - 1), (uint32_t)fn, (uint32_t)(i + l - fn)});
            have = true;
        } else if (have && l) {
            g.residues.insert(g.residues.end(), (const uint8_t*)t + i, (const uint8_t*)t + i + l);
            len += (uint32_t)l;
        }
        i = e + 1;
    }
    flush();
    // a FASTA with fig-style ids keeps only its pegs (SEEDtk types a feature by its fid), else every record
    bool anyPeg = false;
    for (size_t k = 0; k < g.pegs.size() && !anyPeg; k++) anyPeg = g.pegId(k).find(".peg.") != std::string_view::npos;
    if (anyPeg) {
        size_t w = 0, rin = 0, rout = 0;
        for (size_t k = 0; k < g.pegs.size(); k++) {
            const uint32_t L = g.lengths[k];
            if (g.pegId(k).find(".peg.") != std::string_view::npos) {
                if (rout != rin) memmove(g.residues.data() + rout, g.residues.data() + rin, L);
                g.pegs[w] = g.pegs[k]; g.lengths[w] = L; w++;
                rout += L;
            }
            rin += L;
        }
        g.pegs.resize(w); g.lengths.resize(w); g.residues.resize(rout);
    }
}

void loadOne(const std::string& path, PackedGenome& g) {
    g.pegs.clear(); g.lengths.clear(); g.residues.clear(); g.text.clear();
    const size_t slash = path.find_last_of('/');
    const std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
    if (endsWith(base, ".gto")) {
        Genome genome(path);                          // JSON: ids / functions / proteins as strings
        g.id = genome.getId(); g.name = genome.getName();
        for (const Feature& f : genome.getFeatures()) {
            if (!f.isPeg()) continue;                 // Genome.getPegs(): features typed by their fid
            const std::string& prot = f.getProteinTranslation();
            PackedPeg p;
            p.idOff = (uint32_t)g.text.size(); p.idLen = (uint32_t)f.getId().size(); g.text += f.getId();
            p.fnOff = (uint32_t)g.text.size(); p.fnLen = (uint32_t)f.getFunction().size(); g.text += f.getFunction();
            g.pegs.push_back(p);
            g.lengths.push_back((uint32_t)prot.size());
            g.residues.insert(g.residues.end(), prot.begin(), prot.end());
        }
    } else {
        g.text = readFile(path);
        g.id = g.name = base.substr(0, base.find_last_of('.'));
        scanFasta(g);
    }
}

}  // namespace

PackedBatch::~PackedBatch() {
    ka_host_free(codes_); ka_host_free(offsets_); ka_host_free(role_); ka_host_free(hits_); ka_host_free(flag_);
}

void PackedBatch::reserve(uint64_t residues, size_t pegs) {
    const size_t wantCodes = (size_t)((residues * 5 + 7) / 8 + 64);
    if (wantCodes > codesCap_) {
        ka_host_free(codes_);
        codesCap_ = wantCodes + wantCodes / 4;
        codes_ = (uint8_t*)ka_host_alloc(codesCap_);
        if (!codes_) { codesCap_ = 0; throw IOException("Out of pinned host memory for the residue stream."); }
    }
    if (pegs + 1 > pegCap_) {
        ka_host_free(offsets_); ka_host_free(role_); ka_host_free(hits_); ka_host_free(flag_);
        pegCap_ = pegs + 1 + pegs / 4;
        offsets_ = (uint32_t*)ka_host_alloc(pegCap_ * 4);
        role_ = (int32_t*)ka_host_alloc(pegCap_ * 4);
        hits_ = (int32_t*)ka_host_alloc(pegCap_ * 4);
        flag_ = (uint8_t*)ka_host_alloc(pegCap_);
        if (!offsets_ || !role_ || !hits_ || !flag_) { pegCap_ = 0; throw IOException("Out of pinned host memory for the batch."); }
    }
}

void PackedBatch::load(const std::string* files, size_t n, int threads) {
    if (genomes_.size() < n) genomes_.resize(n);
    genomes_.resize(n);
    std::vector<std::string> errors(n);
    const size_t nt = std::min<size_t>((size_t)std::max(1, threads), std::max<size_t>(n, 1));
    auto parallel = [&](auto&& body) {
        std::atomic<size_t> next{0};
        auto work = [&] { for (size_t i = next++; i < n; i = next++) body(i); };
        std::vector<std::thread> th;
        for (size_t t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    };
    const auto tp0 = std::chrono::steady_clock::now();
    // 1. parse: every file into its own staging (bytes) and peg table
    parallel([&](size_t i) {
        try { loadOne(files[i], genomes_[i]); }
        catch (const std::exception& e) { errors[i] = e.what(); if (errors[i].empty()) errors[i] = "unknown error"; }
    });
    for (size_t i = 0; i < n; i++)
        if (!errors[i].empty()) throw IOException("Error loading " + files[i] + ": " + errors[i]);
    // 2. place the genomes in the batch
    nPegs_ = 0; nResidues_ = 0;
    for (PackedGenome& g : genomes_) {
        g.firstPeg = nPegs_; g.firstResidue = nResidues_;
        nPegs_ += g.pegs.size(); nResidues_ += g.residues.size();
    }
    if (nResidues_ >= 0xffffff00ull) throw IOException("A batch holds more than 2^32 residues; lower --batch.");
    reserve(nResidues_, nPegs_);
    const auto tp1 = std::chrono::steady_clock::now();
    // 3. pack: every thread writes the whole stream bytes of its genomes (8 residues = 5 bytes); the group
    //    that straddles two genomes is completed afterwards
    parallel([&](size_t i) {
        PackedGenome& g = genomes_[i];
        uint32_t* off = offsets_ + g.firstPeg;
        uint64_t r = g.firstResidue;
        for (size_t k = 0; k < g.lengths.size(); k++) { off[k] = (uint32_t)r; r += g.lengths[k]; }
        const uint64_t a = g.firstResidue, b = a + g.residues.size();
        const uint64_t a8 = (a + 7) & ~7ull, b8 = b & ~7ull;
        if (a8 < b8) engine_.packResidues(g.residues.data() + (a8 - a), b8 - a8, a8, codes_);
    });
    offsets_[nPegs_] = (uint32_t)nResidues_;
    for (size_t i = 0; i < n; i++) {
        // the (at most two) partial groups at the ends of genome i: gather their 8 residues across genomes
        const PackedGenome& g = genomes_[i];
        const uint64_t a = g.firstResidue, b = a + g.residues.size();
        for (uint64_t grp : {a & ~7ull, b & ~7ull}) {
            if (grp >= nResidues_ || (grp >= ((a + 7) & ~7ull) && grp < (b & ~7ull))) continue;   // inside the bulk part
            uint8_t tmp[8];
            const uint64_t cnt = std::min<uint64_t>(8, nResidues_ - grp);
            size_t j = i;
            while (j > 0 && genomes_[j].firstResidue > grp) j--;
            for (uint64_t q = 0; q < cnt; q++) {
                const uint64_t pos = grp + q;
                while (j + 1 < n && genomes_[j + 1].firstResidue <= pos) j++;
                tmp[q] = genomes_[j].residues[pos - genomes_[j].firstResidue];
            }
            engine_.packResidues(tmp, cnt, grp, codes_);
        }
    }
    parseSeconds = std::chrono::duration<double>(tp1 - tp0).count();
    packSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tp1).count();
}

void PackedBatch::annotate(int minHits) {
    engine_.annotatePacked(codes_, offsets_, nPegs_, minHits, role_, hits_, flag_);
}

}  // namespace theseed
