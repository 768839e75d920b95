// GeneCopyProcessor.hpp — C++ mirror of genome/compare/GeneCopyProcessor.java (`genes`):
//   genes [-m|--maxDist d] [-K|--kmer|--kmerSize n] [--devices i] source.gto target.gto output.gto
// Same lifecycle (setDefaults :82-86, validateParms :88-108, runCommand :110-166), option names,
// defaults (maxDist 0.5, K 8) and messages.  The ProteinKmers comparisons of :137-142 — one k-mer set
// per target peg against the source pegs of the same function — are collected for the whole genome
// and run in ONE ka_kmer_distance call; the selection loop of :139-146 (`f2Dist <= fDist`: the
// closest candidate within maxDist, a later one wins a tie) is then replayed on the distances.
//
// RECALLED, NOT READ (external SEEDtk classes): Genome / Feature JSON layout — features carry
// `alias_pairs` = [[type, alias], …] (Feature.getAliasMap sorts types and aliases; addAlias
// appends a pair that is not there yet); FunctionMap.findOrInsert / getByName identify functions
// after Function normalisation — here: the comment (" # …" / " ! …") is dropped, the text is
// lower-cased and runs of non-alphanumerics collapse to one space.  Genome.save re-serialises
// the JSON: member order and number text are preserved here, whitespace is not.
#pragma once
#include <cctype>
#include <iostream>
#include <string>
#include <vector>

#include "Genome.hpp"
#include "JsonDoc.hpp"
#include "KmerEngine.hpp"

namespace theseed {

/** Function.normalize stand-in (see the header comment). */
inline std::string normalizeFunction(const std::string& funDesc) {
    // drop the comment: first whitespace* followed by '#' or '!' and at least one more character
    std::string s = funDesc;
    for (size_t i = 0; i + 1 < s.size(); i++)
        if (s[i] == '#' || s[i] == '!') {
            size_t b = i;
            while (b > 0 && (s[b - 1] == ' ' || s[b - 1] == '\t')) b--;
            s.erase(b);
            break;
        }
    std::string out;
    bool gap = false;
    for (unsigned char c : s) {
        if (std::isalnum(c)) {
            if (gap && !out.empty()) out.push_back(' ');
            gap = false;
            out.push_back((char)std::tolower(c));
        } else gap = true;
    }
    return out;
}

class GeneCopyProcessor {
public:
    explicit GeneCopyProcessor(std::ostream& log = std::cerr) : log_(log) {}
    bool parseCommand(const std::vector<std::string>& args);
    int run();
    void setDefaults();      // GeneCopyProcessor.java:82-86
    void validateParms();    // :88-108
    void runCommand();       // :110-166
    int getUpdates() const { return updates_; }
    static void usage(std::ostream& os);
private:
    std::ostream& log_;
    double maxDist_ = 0.5;
    int kmerSize_ = 8;
    std::vector<int> devices_{0};
    std::string sourceFile_, targetFile_, outputFile_;
    JsonValue source_, target_;
    int updates_ = 0;
};

}  // namespace theseed
