// GeneCopyProcessor.cpp — see GeneCopyProcessor.hpp.  Line references: genome/compare/GeneCopyProcessor.java.
#include "GeneCopyProcessor.hpp"

#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <unistd.h>
#include <unordered_map>

namespace theseed {

namespace {

bool canRead(const std::string& path) { return access(path.c_str(), R_OK) == 0; }

bool isPeg(const JsonValue& feat) { return feat.getString("id").find(".peg.") != std::string::npos; }

/** Feature.getAliasMap: type -> sorted aliases, from alias_pairs */
std::map<std::string, std::set<std::string>> aliasMapOf(const JsonValue& feat) {
    std::map<std::string, std::set<std::string>> out;
    const JsonValue* pairs = feat.find("alias_pairs");
    if (pairs && pairs->kind == JsonValue::Array)
        for (const JsonValue& pr : pairs->items)
            if (pr.kind == JsonValue::Array && pr.items.size() == 2 && pr.items[0].kind == JsonValue::String &&
                pr.items[1].kind == JsonValue::String)
                out[pr.items[0].text].insert(pr.items[1].text);
    return out;
}

/** Feature.addAlias: append the pair unless the feature already has it */
void addAlias(JsonValue& feat, const std::string& type, const std::string& alias) {
    JsonValue* pairs = feat.find("alias_pairs");
    if (!pairs || pairs->kind != JsonValue::Array) pairs = &feat.set("alias_pairs", JsonValue::array());
    for (const JsonValue& pr : pairs->items)
        if (pr.kind == JsonValue::Array && pr.items.size() == 2 && pr.items[0].text == type && pr.items[1].text == alias) return;
    JsonValue pr = JsonValue::array();
    pr.items.push_back(JsonValue::str(type));
    pr.items.push_back(JsonValue::str(alias));
    pairs->items.push_back(std::move(pr));
}

std::string genomeLabel(const JsonValue& g) { return g.getString("id") + " (" + g.getString("scientific_name") + ")"; }

}  // namespace

void GeneCopyProcessor::usage(std::ostream& os) {
    os << "genes [-m|--maxDist 0.2] [-K|--kmer|--kmerSize 10] [--devices i] source.gto target.gto output.gto\n"
          " -m, --maxDist          maximum permissible distance for a name transfer (default 0.5)\n"
          " -K, --kmer, --kmerSize protein kmer size for distance computation (default 8)\n"
          " --devices              CUDA device of the engine (default 0)\n";
}

void GeneCopyProcessor::setDefaults() {
    maxDist_ = 0.5;       // :84
    kmerSize_ = 8;        // :85
}

bool GeneCopyProcessor::parseCommand(const std::vector<std::string>& args) {
    setDefaults();
    std::vector<std::string> pos;
    try {
        for (size_t i = 0; i < args.size(); i++) {
            const std::string& a = args[i];
            auto value = [&]() -> const std::string& {
                if (i + 1 >= args.size()) throw ParseFailureException("Option \"" + a + "\" takes an operand");
                return args[++i];
            };
            if (a == "-h" || a == "--help") { usage(log_); return false; }
            else if (a == "-m" || a == "--maxDist") {
                char* end = nullptr;
                const std::string& v = value();
                maxDist_ = std::strtod(v.c_str(), &end);
                if (end == v.c_str() || *end) throw ParseFailureException("\"" + v + "\" is not a valid value for \"" + a + "\"");
            }
            else if (a == "-K" || a == "--kmer" || a == "--kmerSize") kmerSize_ = std::atoi(value().c_str());
            else if (a == "--devices") devices_ = {std::atoi(value().c_str())};
            else if (a.size() > 1 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) throw ParseFailureException("\"" + a + "\" is not a valid option");
            else pos.push_back(a);
        }
        if (pos.size() != 3) throw ParseFailureException("Three arguments are required: source.gto target.gto output.gto");
        sourceFile_ = pos[0]; targetFile_ = pos[1]; outputFile_ = pos[2];
        validateParms();
    } catch (const ParseFailureException& e) {
        log_ << e.what() << "\n";
        usage(log_);
        return false;
    } catch (const std::runtime_error& e) {
        log_ << e.what() << "\n";
        return false;
    }
    return true;
}

void GeneCopyProcessor::validateParms() {
    if (maxDist_ < 0.0 || maxDist_ > 1.0)                                                  // :91-92
        throw ParseFailureException("Distance must be between 0 and 1.");
    if (kmerSize_ < 2)                                                                     // :94-95
        throw ParseFailureException("Kmer size must be at least 2.");
    if (!canRead(sourceFile_))                                                             // :98-99
        throw FileNotFoundException("Input genome file " + sourceFile_ + " not found or unreadable.");
    source_ = JsonValue::parse(readFile(sourceFile_));                                     // :100
    if (!canRead(targetFile_))                                                             // :101-102 (the reference re-tests the source file here)
        throw FileNotFoundException("Input genome file " + targetFile_ + " not found or unreadable.");
    target_ = JsonValue::parse(readFile(targetFile_));                                     // :103
}

int GeneCopyProcessor::run() {
    try {
        runCommand();
        return 0;
    } catch (const std::exception& e) {
        log_ << "Error in command: " << e.what() << "\n";
        return 1;
    }
}

void GeneCopyProcessor::runCommand() {
    // source features with aliases, by function (:112-124)
    log_ << "Processing features in " << genomeLabel(source_) << ".\n";
    std::unordered_map<std::string, std::vector<const JsonValue*>> funFeatures;
    std::unordered_map<const JsonValue*, std::map<std::string, std::set<std::string>>> aliasMap;
    const JsonValue* sFeats = source_.find("features");
    if (sFeats && sFeats->kind == JsonValue::Array)
        for (const JsonValue& feat : sFeats->items) {
            if (!isPeg(feat)) continue;
            auto aliases = aliasMapOf(feat);
            if (aliases.empty()) continue;                                                 // :114
            funFeatures[normalizeFunction(feat.getString("function"))].push_back(&feat);   // :116-119
            aliasMap.emplace(&feat, std::move(aliases));                                   // :121
        }
    log_ << aliasMap.size() << " features with aliases, " << funFeatures.size() << " functions found.\n";

    // every target peg with candidates becomes one query group (:127-142)
    std::vector<uint8_t> residues;
    std::vector<uint64_t> offsets{0};
    std::unordered_map<const JsonValue*, uint32_t> seqOf;
    auto intern = [&](const JsonValue* feat) {
        auto it = seqOf.find(feat);
        if (it != seqOf.end()) return it->second;
        const std::string prot = feat->getString("protein_translation");
        residues.insert(residues.end(), prot.begin(), prot.end());
        offsets.push_back(residues.size());
        uint32_t idx = (uint32_t)offsets.size() - 2;
        seqOf.emplace(feat, idx);
        return idx;
    };
    std::vector<uint32_t> querySeq, cand;
    std::vector<uint64_t> groupOff{0};
    std::vector<JsonValue*> queryFeat;
    std::vector<const JsonValue*> candFeat;
    JsonValue* tFeats = target_.find("features");
    if (tFeats && tFeats->kind == JsonValue::Array)
        for (JsonValue& feat : tFeats->items) {
            if (!isPeg(feat)) continue;
            auto it = funFeatures.find(normalizeFunction(feat.getString("function")));     // :129-132
            if (it == funFeatures.end()) continue;
            querySeq.push_back(intern(&feat));
            queryFeat.push_back(&feat);
            for (const JsonValue* f2 : it->second) { cand.push_back(intern(f2)); candFeat.push_back(f2); }
            groupOff.push_back(cand.size());
        }

    std::vector<int32_t> common;
    std::vector<double> dist;
    if (!cand.empty()) {
        KmerEngine engine(devices_);
        engine.kmerDistance(residues, offsets, kmerSize_, querySeq, groupOff, cand, common, dist);   // :137-142
    }

    // the closest candidate within maxDist gives its aliases (:138-160)
    updates_ = 0;
    for (size_t q = 0; q < queryFeat.size(); q++) {
        const JsonValue* found = nullptr;
        double fDist = maxDist_;
        for (uint64_t m = groupOff[q]; m < groupOff[q + 1]; m++)
            if (dist[m] <= fDist) { fDist = dist[m]; found = candFeat[m]; }                // :143-146
        if (found) {
            for (const auto& entry : aliasMap[found])                                      // :153-157
                for (const std::string& alias : entry.second) addAlias(*queryFeat[q], entry.first, alias);
            updates_++;
        }
    }
    log_ << "Writing genome with " << updates_ << " updates to " << outputFile_ << ".\n";   // :164
    std::ofstream out(outputFile_, std::ios::binary);
    if (!out) throw IOException("Cannot write " + outputFile_ + ".");
    out << target_.dump() << "\n";                                                         // :165
}

}  // namespace theseed
