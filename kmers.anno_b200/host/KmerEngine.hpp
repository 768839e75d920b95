// KmerEngine.hpp — RAII C++ face of the C ABI (include/kmeranno.h), the C++ twin of the Java
// `org.theseed.proteins.kmers.gpu.KmerEngine` class shown in INTEGRATION.md.  It replaces the
// `Map<String,String> kmerRoleMap` field of ApplyKmerProcessor.java:53 and the peg loop :122-148.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "kmeranno.h"

namespace theseed {

class KmerEngineError : public std::runtime_error {
public:
    KmerEngineError(int code, const std::string& msg)
        : std::runtime_error("kmeranno error " + std::to_string(code) + ": " + msg), code(code) {}
    int code;
};

class KmerEngine {
public:
    explicit KmerEngine(const std::vector<int>& devices = {0}) {
        int rc = ka_create(devices.data(), (int)devices.size(), &h_);
        if (rc != KA_OK) throw KmerEngineError(rc, ka_last_error(nullptr));
    }
    ~KmerEngine() { ka_destroy(h_); }
    KmerEngine(const KmerEngine&) = delete;
    KmerEngine& operator=(const KmerEngine&) = delete;

    void setOption(const std::string& name, double v) { check(ka_set_option(h_, name.c_str(), v)); }

    /** kmers = n*K residue bytes back to back; roles = dense role ids (one per k-mer line). */
    void loadDb(const std::vector<uint8_t>& kmers, const std::vector<int32_t>& roles, int K) {
        check(ka_db_load(h_, kmers.data(), roles.data(), roles.size(), K));
    }
    ka_db_info dbInfo() { ka_db_info i; check(ka_db_get_info(h_, &i)); return i; }

    /** Annotate a CSR batch; outputs are resized to the number of sequences. */
    void annotate(const std::vector<uint8_t>& residues, const std::vector<uint64_t>& offsets, int minHits,
                  std::vector<int32_t>& role, std::vector<int32_t>& hits, std::vector<uint8_t>& flag) {
        size_t n = offsets.empty() ? 0 : offsets.size() - 1;
        role.resize(n); hits.resize(n); flag.resize(n);
        check(ka_annotate(h_, residues.data(), offsets.data(), n, minHits, role.data(), hits.data(), flag.data()));
    }
    /** The same on the packed form of the batch (ka_annotate_packed): caller-owned, ideally pinned, buffers. */
    void annotatePacked(const uint8_t* codes, const uint32_t* offsets, size_t n, int minHits,
                        int32_t* role, int32_t* hits, uint8_t* flag) {
        check(ka_annotate_packed(h_, codes, offsets, n, minHits, role, hits, flag));
    }
    /** 5-bit codes of residues[0..n) into the stream at residue index `first` (a multiple of 8); thread-safe. */
    void packResidues(const uint8_t* residues, uint64_t n, uint64_t first, uint8_t* codes) {
        check(ka_pack_residues(h_, residues, n, first, codes));
    }
    ka_stats stats() { ka_stats s; check(ka_get_stats(h_, &s)); return s; }

    /** ProteinKmers.distance of every query against its candidates (GeneCopyProcessor.java:137-142):
     *  query q = sequence querySeq[q], candidates cand[groupOff[q] .. groupOff[q+1]). */
    void kmerDistance(const std::vector<uint8_t>& residues, const std::vector<uint64_t>& offsets, int K,
                      const std::vector<uint32_t>& querySeq, const std::vector<uint64_t>& groupOff,
                      const std::vector<uint32_t>& cand, std::vector<int32_t>& common, std::vector<double>& dist) {
        size_t n = offsets.empty() ? 0 : offsets.size() - 1;
        common.resize(cand.size()); dist.resize(cand.size());
        check(ka_kmer_distance(h_, residues.data(), offsets.data(), n, K, querySeq.data(), groupOff.data(),
                               querySeq.size(), cand.data(), nullptr, common.data(), dist.data()));
    }

private:
    void check(int rc) { if (rc != KA_OK) throw KmerEngineError(rc, ka_last_error(h_)); }
    ka_engine* h_ = nullptr;
};

}  // namespace theseed
