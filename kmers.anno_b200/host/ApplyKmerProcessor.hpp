// ApplyKmerProcessor.hpp — C++ mirror of proteins/kmers/anno/ApplyKmerProcessor.java
// (`apply` command), same lifecycle (setDefaults -> validateParms -> runCommand), same
// options and positional parameters, same error behaviour; the peg loop (:122-148) runs on
// the GPU engine and the reporter calls are replayed in the original peg order.
//
//   apply [--format VERIFY|APPLY] [-m|--min N] kmerdb.tbl roles.in.use gtoDir
// added knobs (ordinary options, as SURVEY §5 asks): --devices 0,1,..  --table-mode 0|1|2  --batch genomes  --threads n
// Genomes of the next batch are parsed and packed into pinned memory by a thread pool (PackedBatch) while the
// GPU annotates the current one.
#pragma once
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "ApplyKmerReporter.hpp"
#include "Genome.hpp"
#include "KmerEngine.hpp"
#include "PackedBatch.hpp"

namespace theseed {

class ApplyKmerProcessor {
public:
    explicit ApplyKmerProcessor(std::ostream& out = std::cout, std::ostream& log = std::cerr)
        : out_(out), log_(log) {}

    /** args4j-style parse; returns false (after printing usage) on -h or a bad command line. */
    bool parseCommand(const std::vector<std::string>& args);
    /** runCommand with the reference's exception -> message behaviour; returns the exit code. */
    int run();

    // reference lifecycle, public for the tests
    void setDefaults();                       // ApplyKmerProcessor.java:76-80
    void validateParms();                     // :82-111
    void runCommand();                        // :113-155

    int getKmerSize() const { return kmerSize_; }
    static void usage(std::ostream& os);

private:
    void flushBatch(PackedBatch& batch);

    std::ostream& out_;
    std::ostream& log_;
    // command-line options (:58-74)
    ApplyKmerReporter::Type outputType_ = ApplyKmerReporter::Type::APPLY;
    int minHits_ = 5;
    std::string kmerDbFile_, goodRoleFile_, inDir_;
    std::vector<int> devices_{0};
    int batchGenomes_ = 32;
    int loadThreads_ = 1;
    int tableMode_ = 0;                        // 0 replicated table, 1 / 2 sharded over the devices (ka_set_option "table_mode")
    // state
    std::unique_ptr<ApplyKmerReporter> reporter_;
    std::unique_ptr<KmerEngine> engine_;       // replaces Map<String,String> kmerRoleMap (:53)
    std::vector<std::string> roleNames_;       // dense role id -> role string
    std::vector<int> roleColumn_;              // dense role id -> report column (getRoleIdx), 0 = not in roles.in.use
    uint64_t proteinsDone_ = 0, residuesDone_ = 0;
    int kmerSize_ = 0;
};

}  // namespace theseed
