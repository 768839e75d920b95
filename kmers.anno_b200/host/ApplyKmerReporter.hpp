// ApplyKmerReporter.hpp — C++ mirror of the reference's report classes, method for method:
//   reports/ApplyKmerReporter.java:21-126        (role -> column map, abstract hooks, Type enum)
//   reports/DefaultApplyKmerReporter.java:17-61  (APPLY: one row of per-role peg counts per genome)
//   reports/VerifyApplyKmerReporter.java:13-55   (VERIFY: one row per called peg)
#pragma once
#include <cstdio>
#include <memory>
#include <ostream>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "Genome.hpp"

namespace theseed {

class ApplyKmerReporter {
public:
    enum class Type { VERIFY, APPLY };                       // ApplyKmerReporter.java:107-108
    static std::unique_ptr<ApplyKmerReporter> create(Type t, std::ostream& out);  // :113-124
    static Type parseType(const std::string& s);

    explicit ApplyKmerReporter(std::ostream& out) : out_(out) {}
    virtual ~ApplyKmerReporter() = default;

    /** :43-54 — line n (1-based) of rolesToUse, text before the first TAB, is the role of column n. */
    void initReport(const std::string& rolesToUse) {
        int idx = 1;
        for (const std::string& line : readLines(rolesToUse)) {
            roleIdxMap_[line.substr(0, line.find('\t'))] = idx;   // put(): a repeated role keeps the last index
            idx++;
        }
        openReport();
    }
    virtual void openGenome(const Genome& genome) { openGenome(genome.getId()); }        // :66
    virtual void openGenome(const std::string& genomeId) = 0;
    /** recordFeature (:75) without a Feature object: the ingest pipeline keeps peg ids and functions as views
     *  into the genome file, and knows the report column of every role id (getRoleIdx, computed once). */
    virtual void recordCall(std::string_view pegId, std::string_view function, const std::string& role, int roleIdx, int count) = 0;
    virtual void recordFeature(const Feature& feat, const std::string& role, int count) = 0;  // :75
    virtual void closeGenome() = 0;                                                      // :80
    virtual void closeReport() = 0;                                                      // :85
    void close() { out_.flush(); }

    /** :92-95 — column index of a role, 0 if the role is not interesting */
    int getRoleIdx(const std::string& roleId) const {
        auto it = roleIdxMap_.find(roleId);
        return it == roleIdxMap_.end() ? 0 : it->second;
    }
    /** :100-102 */
    int getNumRoles() const { return (int)roleIdxMap_.size(); }

protected:
    virtual void openReport() = 0;                                                       // :59
    void println(const std::string& s) { out_ << s << '\n'; }
    std::ostream& out_;

private:
    std::unordered_map<std::string, int> roleIdxMap_;
};

class DefaultApplyKmerReporter : public ApplyKmerReporter {
public:
    using ApplyKmerReporter::ApplyKmerReporter;
    void openReport() override { roleCounts_.assign((size_t)getNumRoles(), 0); }          // :33-35
    using ApplyKmerReporter::openGenome;
    void openGenome(const std::string& genomeId) override {                               // :38-41
        genomeId_ = genomeId;
        std::fill(roleCounts_.begin(), roleCounts_.end(), 0);
    }
    void recordFeature(const Feature&, const std::string& role, int) override {           // :44-48
        int idx = getRoleIdx(role);
        if (idx > 0) roleCounts_[(size_t)idx - 1]++;
    }
    void recordCall(std::string_view, std::string_view, const std::string&, int roleIdx, int) override {
        if (roleIdx > 0) roleCounts_[(size_t)roleIdx - 1]++;
    }
    void closeGenome() override {                                                         // :51-55
        line_.assign(genomeId_);
        line_ += '\t';
        char buf[16];
        for (size_t i = 0; i < roleCounts_.size(); i++) {
            if (i) line_ += '\t';
            int v = roleCounts_[i], n = 0;
            if (v < 10) { line_ += (char)('0' + v); continue; }
            while (v) { buf[n++] = (char)('0' + v % 10); v /= 10; }
            while (n) line_ += buf[--n];
        }
        println(line_);
    }
    void closeReport() override {}
private:
    std::vector<int> roleCounts_;
    std::string genomeId_, line_;
};

class VerifyApplyKmerReporter : public ApplyKmerReporter {
public:
    using ApplyKmerReporter::ApplyKmerReporter;
    void openReport() override { println("genome_id\tpeg_id\trole\thits\tfunction"); }    // :33-35
    using ApplyKmerReporter::openGenome;
    void openGenome(const std::string& genomeId) override { genomeId_ = genomeId; }       // :38-40
    void recordFeature(const Feature& feat, const std::string& role, int count) override {  // :43-45
        println(genomeId_ + "\t" + feat.getId() + "\t" + role + "\t" + std::to_string(count) + "\t" + feat.getFunction());
    }
    void recordCall(std::string_view pegId, std::string_view function, const std::string& role, int, int count) override {
        std::string line = genomeId_;
        line += '\t'; line.append(pegId); line += '\t'; line += role; line += '\t'; line += std::to_string(count);
        line += '\t'; line.append(function);
        println(line);
    }
    void closeGenome() override {}
    void closeReport() override {}
private:
    std::string genomeId_;
};

inline std::unique_ptr<ApplyKmerReporter> ApplyKmerReporter::create(Type t, std::ostream& out) {
    if (t == Type::VERIFY) return std::make_unique<VerifyApplyKmerReporter>(out);
    return std::make_unique<DefaultApplyKmerReporter>(out);
}

inline ApplyKmerReporter::Type ApplyKmerReporter::parseType(const std::string& s) {
    if (s == "VERIFY") return Type::VERIFY;
    if (s == "APPLY") return Type::APPLY;
    throw ParseFailureException("\"" + s + "\" is not a valid value for \"--format\" (VERIFY, APPLY).");
}

}  // namespace theseed
