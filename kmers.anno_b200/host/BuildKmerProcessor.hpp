// BuildKmerProcessor.hpp — C++ mirror of proteins/kmers/anno/BuildKmerProcessor.java (`build`):
//   build [-g genomeFile.tbl] [-K n] roles.in.subsystems roles.to.use genomeDir  > kmerdb.tbl
// Same lifecycle and option names; the two passes over the k-mer map (:148-208) run on the GPU
// (ka_build).  Output lines `kmer TAB roleId` (:212-216); their ORDER is unspecified here (the
// reference prints in HashMap iteration order).
//
// RECALLED, NOT READ: org.theseed.proteins.RoleMap / Role / Feature.getUsefulRoles are external
// classes.  This mirror loads `roles.in.subsystems` as id TAB [checksum TAB] name, splits a
// function into roles on " / ", " @ ", "; " after dropping a trailing "#"/"!" comment, and
// matches role names after normalisation (lower case, EC/TC numbers removed, runs of
// non-alphanumerics collapsed) — the documented intent of Role.checksum, not its exact code.
// Unlike the Java option -K (KmerReference.setKmerSize, which never reaches ProteinKmers —
// SURVEY §8a2), -K here really sets the k-mer length; the default is 8 either way.
#pragma once
#include <iostream>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "Genome.hpp"
#include "KmerEngine.hpp"

namespace theseed {

class RoleMap {
public:
    static RoleMap load(const std::string& file);
    /** role id of a role description, "" if the role is not in the map */
    std::string getByName(const std::string& roleDesc) const;
    std::string getName(const std::string& id) const;
    static std::string normalize(const std::string& roleDesc);
    size_t size() const { return byId_.size(); }
private:
    std::unordered_map<std::string, std::string> byNorm_;  // normalised name -> id
    std::unordered_map<std::string, std::string> byId_;    // id -> name
};

/** Feature.rolesOfFunction */
std::vector<std::string> rolesOfFunction(const std::string& function);

class BuildKmerProcessor {
public:
    explicit BuildKmerProcessor(std::ostream& out = std::cout, std::ostream& log = std::cerr) : out_(out), log_(log) {}
    bool parseCommand(const std::vector<std::string>& args);
    int run();
    void setDefaults();      // BuildKmerProcessor.java:101-107
    void validateParms();    // :109-135
    void runCommand();       // :137-223
    static void usage(std::ostream& os);
private:
    std::ostream& out_;
    std::ostream& log_;
    std::string genomeFile_, roleMapFile_, roleIdFile_, gtoDir_;
    int kmerSize_ = 8;
    std::vector<int> devices_{0};
    RoleMap roleMap_;
    std::set<std::string> goodRoles_, goodGenomes_;
    bool filterGenomes_ = false;
};

}  // namespace theseed
