// PackedBatch.hpp — the ingest side of `apply` (ApplyKmerProcessor.java:116-123): the genomes of one
// GPU batch are parsed straight into the engine's packed input form — 5-bit residue codes and
// 32-bit offsets in PINNED host memory (ka_host_alloc) — by a pool of threads, one file at a
// time per thread, with no per-protein std::string and no staging copy; two PackedBatch
// objects alternate so that batch i+1 is parsed and packed while batch i is on the GPU.
//
// FASTA files are scanned in place (the file text is kept for the report's peg ids and
// functions); GTO (JSON) files go through the Genome class.  Peg order is file order (:122).
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>

#include "Genome.hpp"
#include "KmerEngine.hpp"

namespace theseed {

struct PackedPeg {
    uint32_t idOff, idLen, fnOff, fnLen;   // peg id and function inside PackedGenome::text
};

struct PackedGenome {
    std::string id, name;
    std::string text;                      // FASTA: the file; GTO: ids and functions back to back
    std::vector<PackedPeg> pegs;
    std::vector<uint8_t> residues;         // parse-time staging of this file's residues (reused)
    std::vector<uint32_t> lengths;
    size_t firstPeg = 0;                   // index of its first peg in the batch
    uint64_t firstResidue = 0;             // index of its first residue in the batch
    std::string toString() const { return id + " (" + name + ")"; }
    std::string_view pegId(size_t i) const { return std::string_view(text).substr(pegs[i].idOff, pegs[i].idLen); }
    std::string_view pegFunction(size_t i) const { return std::string_view(text).substr(pegs[i].fnOff, pegs[i].fnLen); }
};

class PackedBatch {
public:
    explicit PackedBatch(KmerEngine& engine) : engine_(engine) {}
    ~PackedBatch();
    PackedBatch(const PackedBatch&) = delete;
    PackedBatch& operator=(const PackedBatch&) = delete;

    /** Parse and pack files[0..n) with up to `threads` threads; throws IOException naming the bad file. */
    void load(const std::string* files, size_t n, int threads);
    /** Size the pinned buffers ahead of the first load (cudaHostAlloc is slow: tens of ms per 100 MB). */
    void reserve(uint64_t residues, size_t pegs);
    /** ka_annotate_packed on the whole batch; results in role() / hits() / flag(). */
    void annotate(int minHits);

    size_t numGenomes() const { return genomes_.size(); }
    size_t numPegs() const { return nPegs_; }
    uint64_t numResidues() const { return nResidues_; }
    const PackedGenome& genome(size_t g) const { return genomes_[g]; }
    const int32_t* role() const { return role_; }
    const int32_t* hits() const { return hits_; }
    const uint8_t* flag() const { return flag_; }

private:
    KmerEngine& engine_;
    std::vector<PackedGenome> genomes_;
    size_t nPegs_ = 0;
public:
    double parseSeconds = 0, packSeconds = 0;   // of the last load()
private:
    uint64_t nResidues_ = 0;
    // pinned buffers, grown on demand and reused from batch to batch
    uint8_t* codes_ = nullptr; size_t codesCap_ = 0;
    uint32_t* offsets_ = nullptr; size_t pegCap_ = 0;
    int32_t* role_ = nullptr; int32_t* hits_ = nullptr; uint8_t* flag_ = nullptr;
};

}  // namespace theseed
