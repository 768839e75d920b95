// BuildKmerProcessor.cpp — see BuildKmerProcessor.hpp.  Line references are to
// /root/reference/src/main/java/org/theseed/proteins/kmers/anno/BuildKmerProcessor.java.
#include "BuildKmerProcessor.hpp"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sys/stat.h>

namespace theseed {

namespace {
bool isDirectory(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
bool canRead(const std::string& p) { std::ifstream in(p); return (bool)in && !isDirectory(p); }
std::vector<std::string> splitTabs(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0;
    for (;;) {
        size_t e = line.find('\t', i);
        if (e == std::string::npos) { out.emplace_back(line, i); break; }
        out.emplace_back(line, i, e - i);
        i = e + 1;
    }
    return out;
}
std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
}  // namespace

std::string RoleMap::normalize(const std::string& roleDesc) {
    // drop "(EC 1.2.3.4)" / "(TC 3.A.1.-)" groups, lower-case, collapse everything that is not a
    // letter or digit into one blank
    std::string s;
    for (size_t i = 0; i < roleDesc.size();) {
        if (roleDesc[i] == '(' && i + 3 < roleDesc.size() && (roleDesc.compare(i + 1, 3, "EC ") == 0 || roleDesc.compare(i + 1, 3, "TC ") == 0)) {
            size_t e = roleDesc.find(')', i);
            if (e != std::string::npos) { i = e + 1; continue; }
        }
        s.push_back(roleDesc[i++]);
    }
    std::string out;
    bool blank = true;
    for (unsigned char c : s) {
        if (std::isalnum(c)) { out.push_back((char)std::tolower(c)); blank = false; }
        else if (!blank) { out.push_back(' '); blank = true; }
    }
    while (!out.empty() && out.back() == ' ') out.pop_back();
    return out;
}

RoleMap RoleMap::load(const std::string& file) {
    RoleMap m;
    for (const std::string& line : readLines(file)) {
        if (line.empty()) continue;
        std::vector<std::string> f = splitTabs(line);
        if (f.size() < 2) continue;
        const std::string& id = f[0];
        const std::string& name = f.back();          // id TAB name, or id TAB checksum TAB name
        m.byId_[id] = name;
        m.byNorm_[normalize(name)] = id;
    }
    return m;
}

std::string RoleMap::getByName(const std::string& roleDesc) const {
    auto it = byNorm_.find(normalize(roleDesc));
    return it == byNorm_.end() ? "" : it->second;
}

std::string RoleMap::getName(const std::string& id) const {
    auto it = byId_.find(id);
    return it == byId_.end() ? "" : it->second;
}

std::vector<std::string> rolesOfFunction(const std::string& function) {
    std::string fun = function;
    size_t c = fun.find_first_of("#!");
    if (c != std::string::npos) fun = fun.substr(0, c);
    std::vector<std::string> out;
    size_t i = 0;
    while (i <= fun.size()) {
        size_t best = std::string::npos, blen = 0;
        for (const char* sep : {" / ", " @ ", "; "}) {
            size_t p = fun.find(sep, i);
            if (p < best) { best = p; blen = std::string(sep).size(); }
        }
        std::string part = trim(best == std::string::npos ? fun.substr(i) : fun.substr(i, best - i));
        if (!part.empty()) out.push_back(part);
        if (best == std::string::npos) break;
        i = best + blen;
    }
    return out;
}

void BuildKmerProcessor::usage(std::ostream& os) {
    os << "build [-g genomeFile.tbl] [-K n] [--devices 0,..] roles.in.subsystems roles.to.use genomeDir\n"
          " roles.in.subsystems  role definition file\n roles.to.use         interesting role file\n"
          " genomeDir            input genome directory\n -g (--genomes)       file of acceptable genome IDs\n"
          " -K (--kmer)          protein kmer length (default 8)\n";
}

void BuildKmerProcessor::setDefaults() {
    genomeFile_.clear();            // :103
    filterGenomes_ = false;         // goodGenomes = null :105
    kmerSize_ = 8;                  // KmerReference.setKmerSize(8) :106
    devices_ = {0};
}

bool BuildKmerProcessor::parseCommand(const std::vector<std::string>& args) {
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
            else if (a == "-g" || a == "--genomes") genomeFile_ = value();
            else if (a == "-K" || a == "--kmer") kmerSize_ = std::atoi(value().c_str());
            else if (a == "-t" || a == "--workDir") value();   // no temporary FASTA buffer is needed here
            else if (a == "--devices") devices_ = {std::atoi(value().c_str())};
            else if (a.size() > 1 && a[0] == '-') throw ParseFailureException("\"" + a + "\" is not a valid option");
            else pos.push_back(a);
        }
        if (pos.size() != 3) throw ParseFailureException("Three arguments are required: roles.in.subsystems roles.to.use genomeDir");
        roleMapFile_ = pos[0]; roleIdFile_ = pos[1]; gtoDir_ = pos[2];
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

void BuildKmerProcessor::validateParms() {
    if (genomeFile_.empty()) log_ << "No genome filtering.\n";                                   // :112-113
    else {
        if (!canRead(genomeFile_)) throw FileNotFoundException("Good-genome file " + genomeFile_ + " not found or unreadable.");
        std::vector<std::string> lines = readLines(genomeFile_);                                 // TabbedLineReader.readSet(file, "1")
        for (size_t i = 1; i < lines.size(); i++) if (!lines[i].empty()) goodGenomes_.insert(splitTabs(lines[i])[0]);
        filterGenomes_ = true;
        log_ << goodGenomes_.size() << " genome IDs read from genome-filter file.\n";
    }
    log_ << "Reading role definitions from " << roleMapFile_ << ".\n";                          // :121-122
    roleMap_ = RoleMap::load(roleMapFile_);
    if (!canRead(roleIdFile_)) throw FileNotFoundException("Good-role file " + roleIdFile_ + " not found or unreadable.");
    for (const std::string& line : readLines(roleIdFile_))                                       // LineReader.readSet :126
        if (!line.empty()) goodRoles_.insert(splitTabs(line)[0]);
    if (!isDirectory(gtoDir_)) throw FileNotFoundException("Genome directory " + gtoDir_ + " not found or invalid.");   // :133-134
    if (kmerSize_ < 1 || kmerSize_ > 12) throw ParseFailureException("Kmer length must be between 1 and 12 for the GPU engine.");
}

void BuildKmerProcessor::runCommand() {
    GenomeDirectory genomes(gtoDir_);                                                            // :146
    log_ << genomes.size() << " genomes to scan.\n";
    std::vector<uint8_t> residues;
    std::vector<uint64_t> offsets{0};
    std::vector<int32_t> nRoles, pegRole;
    std::vector<std::string> roleIds;
    std::unordered_map<std::string, int32_t> roleIdx;
    for (const std::string& file : genomes.files()) {
        Genome genome(file);
        if (filterGenomes_ && !goodGenomes_.count(genome.getId())) continue;                     // :149
        log_ << "Processing " << genome.toString() << ".\n";
        int buffered = 0, interesting = 0;
        for (const Feature* peg : genome.getPegs()) {                                            // :157
            std::vector<std::string> pegRoles;                                                   // getUsefulRoles ∩ goodRoles :158
            for (const std::string& desc : rolesOfFunction(peg->getFunction())) {
                std::string id = roleMap_.getByName(desc);
                if (!id.empty() && goodRoles_.count(id)) pegRoles.push_back(id);
            }
            const std::string& prot = peg->getProteinTranslation();
            residues.insert(residues.end(), prot.begin(), prot.end());
            offsets.push_back(residues.size());
            nRoles.push_back((int32_t)pegRoles.size());
            int32_t r = -1;
            if (pegRoles.size() == 1) {                                                          // :165
                auto it = roleIdx.find(pegRoles[0]);
                if (it == roleIdx.end()) { it = roleIdx.emplace(pegRoles[0], (int32_t)roleIds.size()).first; roleIds.push_back(pegRoles[0]); }
                r = it->second;
                interesting++;
            } else if (pegRoles.empty()) buffered++;                                             // :159-164
            pegRole.push_back(r);
        }
        log_ << interesting << " interesting pegs found, " << buffered << " buffered.\n";        // :177
    }
    uint64_t cap = 0;
    for (size_t i = 0; i < nRoles.size(); i++) {
        uint64_t L = offsets[i + 1] - offsets[i];
        if (nRoles[i] == 1 && L >= (uint64_t)kmerSize_) cap += L - kmerSize_ + 1;
    }
    std::vector<uint8_t> kmers(std::max<uint64_t>(cap, 1) * kmerSize_);
    std::vector<int32_t> roles(std::max<uint64_t>(cap, 1));
    uint64_t n = 0;
    ka_engine* e = nullptr;
    int rc = ka_create(devices_.data(), (int)devices_.size(), &e);
    if (rc != KA_OK) throw KmerEngineError(rc, ka_last_error(nullptr));
    rc = ka_build(e, residues.data(), offsets.data(), nRoles.size(), nRoles.data(), pegRole.data(), kmerSize_, cap,
                  kmers.data(), roles.data(), &n, 0);                                            // :148-208 on the GPU
    std::string err = rc != KA_OK ? ka_last_error(e) : "";
    ka_destroy(e);
    if (rc != KA_OK) throw KmerEngineError(rc, err);
    log_ << n << " discriminating kmers remaining.\n";                                           // :209
    std::vector<uint64_t> perRole(roleIds.size(), 0);
    std::string line;
    for (uint64_t i = 0; i < n; i++) {                                                           // :212-216
        line.assign(kmers.begin() + i * kmerSize_, kmers.begin() + (i + 1) * kmerSize_);
        out_ << line << '\t' << roleIds[(size_t)roles[i]] << '\n';
        perRole[(size_t)roles[i]]++;
    }
    out_.flush();
    for (const std::string& id : goodRoles_) {                                                   // :217-222
        auto it = roleIdx.find(id);
        if (it == roleIdx.end() || perRole[(size_t)it->second] == 0)
            log_ << "No kmers found for " << id << ": " << roleMap_.getName(id) << ".\n";
    }
}

int BuildKmerProcessor::run() {
    try { runCommand(); return 0; }
    catch (const std::exception& e) { log_ << "EXECUTION ERROR: " << e.what() << "\n"; return 1; }
}

}  // namespace theseed
