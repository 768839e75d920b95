// ApplyKmerProcessor.cpp — see ApplyKmerProcessor.hpp.  Line references are to
// /root/reference/src/main/java/org/theseed/proteins/kmers/anno/ApplyKmerProcessor.java.
#include "ApplyKmerProcessor.hpp"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <future>
#include <sys/stat.h>
#include <thread>

namespace theseed {

namespace {
double nowSeconds() {
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
bool isDirectory(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool canRead(const std::string& p) {
    std::ifstream in(p);
    return (bool)in && !isDirectory(p);
}
int parseInt(const std::string& opt, const std::string& v) {
    char* end = nullptr;
    long x = std::strtol(v.c_str(), &end, 10);
    if (v.empty() || *end) throw ParseFailureException("\"" + v + "\" is not a valid value for \"" + opt + "\"");
    return (int)x;
}
}  // namespace

void ApplyKmerProcessor::usage(std::ostream& os) {
    os << "apply [--format VERIFY|APPLY] [-m|--min N] [--devices 0,1,..] [--table-mode 0|1|2|3] [--batch N] kmerdb.tbl roles.in.use gtoDir\n"
          " kmerdb.tbl     discriminating kmer database\n"
          " roles.in.use   list of roles in use\n"
          " gtoDir         input genome directory\n"
          " --format       reporting format (default APPLY)\n"
          " -m (--min)     minimum number of hits required to call a role (default 5)\n"
          " --devices      CUDA devices to shard the sequences over (default 0)\n"
          " --batch        genomes per GPU batch (default 32)\n"
          " --threads      genome-parsing threads (default: hardware threads)\n";
}

void ApplyKmerProcessor::setDefaults() {
    outputType_ = ApplyKmerReporter::Type::APPLY;  // :78
    minHits_ = 5;                                  // :79
    devices_ = {0};
    batchGenomes_ = 32;
    loadThreads_ = (int)std::max(1u, std::thread::hardware_concurrency());
}

bool ApplyKmerProcessor::parseCommand(const std::vector<std::string>& args) {
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
            else if (a == "--format") outputType_ = ApplyKmerReporter::parseType(value());
            else if (a == "-m" || a == "--min") minHits_ = parseInt(a, value());
            else if (a == "--batch") batchGenomes_ = parseInt(a, value());
            else if (a == "--threads") loadThreads_ = parseInt(a, value());
            else if (a == "--table-mode") {
                tableMode_ = parseInt(a, value());
                if (tableMode_ < 0 || tableMode_ > 3) throw ParseFailureException("--table-mode must be 0 (replicated), 1 (sharded, peer loads), 2 (sharded, NCCL routed) or 3 (sharded, routed by peer stores).");
            }
            else if (a == "--devices") {
                devices_.clear();
                const std::string& v = value();
                size_t s = 0;
                while (s <= v.size()) {
                    size_t e = v.find(',', s);
                    if (e == std::string::npos) e = v.size();
                    devices_.push_back(parseInt(a, v.substr(s, e - s)));
                    s = e + 1;
                }
            } else if (a.size() > 1 && a[0] == '-') throw ParseFailureException("\"" + a + "\" is not a valid option");
            else pos.push_back(a);
        }
        if (pos.size() < 3) throw ParseFailureException("Argument \"" + std::string(pos.empty() ? "kmerdb.tbl" : pos.size() == 1 ? "roles.in.use" : "gtoDir") + "\" is required");
        if (pos.size() > 3) throw ParseFailureException("Too many arguments: " + pos[3]);
        kmerDbFile_ = pos[0]; goodRoleFile_ = pos[1]; inDir_ = pos[2];
        validateParms();
    } catch (const ParseFailureException& e) {
        log_ << e.what() << "\n";
        usage(log_);
        return false;
    } catch (const std::runtime_error& e) {   // FileNotFoundException / IOException / engine errors
        log_ << e.what() << "\n";
        return false;
    }
    return true;
}

void ApplyKmerProcessor::validateParms() {
    // Verify the input directory.  (:85-86)
    if (!isDirectory(inDir_)) throw FileNotFoundException("Input directory " + inDir_ + " not found or invalid.");
    // Verify the kmer database.  (:88-89)
    if (!canRead(kmerDbFile_)) throw FileNotFoundException("Kmer database file " + kmerDbFile_ + " not found or unreadable.");
    // Verify the minimum number of hits.  (:91-92)
    if (minHits_ < 1) throw ParseFailureException("Min-hits must be positive.");
    if (batchGenomes_ < 1) throw ParseFailureException("Batch size must be positive.");
    // Initialize the reporting.  (:94-98)
    reporter_ = ApplyKmerReporter::create(outputType_, out_);
    if (!canRead(goodRoleFile_)) throw FileNotFoundException("Roles-to-use file " + goodRoleFile_ + " not found or unreadable.");
    log_ << "Reading roles to use from " << goodRoleFile_ << ".\n";
    reporter_->initReport(goodRoleFile_);
    // Load the kmer database.  (:100-110)  TabbedLineReader(file, 2): headerless, two columns.
    log_ << "Loading kmer database from " << kmerDbFile_ << ".  (t=" << nowSeconds() << " s)\n";
    std::string text = readFile(kmerDbFile_);
    // The file is cut at line boundaries into one slice per thread; every slice is parsed into
    // its own k-mer bytes / local role ids, then the role strings are interned globally (the ids
    // are internal: any consistent numbering gives the same report).
    struct Slice {
        size_t begin = 0, end = 0, firstLine = 0;
        std::vector<uint8_t> kmers;
        std::vector<int32_t> roles;
        std::vector<std::string> names;
        std::unordered_map<std::string, int32_t> ids;
        int K = -1;
        std::string error;
    };
    size_t nt = std::max<size_t>(1, std::min<size_t>((size_t)loadThreads_, text.size() / (1 << 20) + 1));
    std::vector<Slice> slices(nt);
    for (size_t t = 0; t < nt; t++) {
        size_t b0 = text.size() * t / nt;
        if (t && b0 < text.size()) { b0 = text.find('\n', b0); b0 = b0 == std::string::npos ? text.size() : b0 + 1; }
        slices[t].begin = t ? b0 : 0;
        if (t) slices[t - 1].end = slices[t].begin;
    }
    slices[nt - 1].end = text.size();
    auto parse = [&](Slice& sl) {
        sl.kmers.reserve((sl.end - sl.begin) / 2);
        sl.roles.reserve((sl.end - sl.begin) / 30 + 16);   // the reference's own sizing hint (:101)
        size_t i = sl.begin, lineNo = 0;
        std::string role;
        while (i < sl.end) {
            size_t e = text.find('\n', i);
            if (e == std::string::npos || e > sl.end) e = sl.end;
            size_t len = e - i;
            if (len && text[i + len - 1] == '\r') len--;
            lineNo++;
            if (len) {
                size_t tab = text.find('\t', i);
                if (tab == std::string::npos || tab >= i + len) { sl.error = "a line has fewer than 2 columns"; sl.firstLine = lineNo; return; }
                size_t klen = tab - i;
                size_t rend = text.find('\t', tab + 1);
                if (rend == std::string::npos || rend > i + len) rend = i + len;
                // HashMap<String,String> takes k-mers of any length; one packed table needs one K
                if (sl.K < 0) sl.K = (int)klen;
                else if ((int)klen != sl.K) { sl.error = "mixed k-mer lengths (" + std::to_string(sl.K) + " and " + std::to_string(klen) + ")"; sl.firstLine = lineNo; return; }
                sl.kmers.insert(sl.kmers.end(), text.begin() + i, text.begin() + tab);
                role.assign(text, tab + 1, rend - tab - 1);
                auto it = sl.ids.find(role);
                if (it == sl.ids.end()) { it = sl.ids.emplace(role, (int32_t)sl.names.size()).first; sl.names.push_back(role); }
                sl.roles.push_back(it->second);
            }
            i = e + 1;
        }
    };
    {
        std::vector<std::thread> th;
        for (size_t t = 1; t < nt; t++) th.emplace_back(parse, std::ref(slices[t]));
        parse(slices[0]);
        for (auto& x : th) x.join();
    }
    std::vector<uint8_t> kmers;
    std::vector<int32_t> roles;
    std::unordered_map<std::string, int32_t> roleIds;
    int K = -1;
    for (Slice& sl : slices) {
        if (!sl.error.empty()) throw IOException("Kmer database " + kmerDbFile_ + ": " + sl.error + " (not supported by the GPU engine).");
        if (sl.K < 0) continue;
        if (K < 0) K = sl.K;
        else if (sl.K != K) throw IOException("Kmer database " + kmerDbFile_ + " mixes k-mer lengths (" + std::to_string(K) + " and " + std::to_string(sl.K) + "): not supported by the GPU engine.");
        std::vector<int32_t> remap(sl.names.size());
        for (size_t j = 0; j < sl.names.size(); j++) {
            auto it = roleIds.find(sl.names[j]);
            if (it == roleIds.end()) { it = roleIds.emplace(sl.names[j], (int32_t)roleNames_.size()).first; roleNames_.push_back(sl.names[j]); }
            remap[j] = it->second;
        }
        kmers.insert(kmers.end(), sl.kmers.begin(), sl.kmers.end());
        for (int32_t r : sl.roles) roles.push_back(remap[(size_t)r]);
        sl.kmers.clear(); sl.kmers.shrink_to_fit(); sl.roles.clear(); sl.roles.shrink_to_fit();
    }
    if (roles.empty()) throw IOException("Kmer database " + kmerDbFile_ + " is empty.");
    kmerSize_ = K;                               // KmerReference.setKmerSize(kmer.length()) (:108)
    log_ << "Kmer size is " << kmerSize_ << ".  (" << roles.size() << " lines parsed, t=" << nowSeconds() << " s)\n";
    engine_ = std::make_unique<KmerEngine>(devices_);   // throws if there is no usable GPU: no CPU fallback
    if (tableMode_) engine_->setOption("table_mode", tableMode_);   // table beyond one GPU: sharded over --devices
    engine_->loadDb(kmers, roles, K);
    ka_db_info info = engine_->dbInfo();
    log_ << info.n_keys << " distinct kmers for " << roleNames_.size() << " roles loaded on " << devices_.size()
         << " device(s), " << info.table_bytes / (1024 * 1024) << " MiB table, " << info.slot_bits << "-bit slots.  (t=" << nowSeconds() << " s)\n";
}

void ApplyKmerProcessor::flushBatch(PackedBatch& batch) {
    if (batch.numGenomes() == 0) return;
    batch.annotate(minHits_);                                             // the peg loop :122-148 for the whole batch
    const int32_t* role = batch.role();
    const int32_t* hits = batch.hits();
    const uint8_t* flag = batch.flag();
    for (size_t g = 0; g < batch.numGenomes(); g++) {
        const PackedGenome& genome = batch.genome(g);
        log_ << "Processing genome " << genome.toString() << ".\n";       // :119
        reporter_->openGenome(genome.id);                                 // :120
        const size_t s0 = genome.firstPeg;
        for (size_t k = 0; k < genome.pegs.size(); k++) {
            const size_t s = s0 + k;
            if (flag[s] == KA_FLAG_CALLED) {                              // :146
                const size_t r = (size_t)role[s];
                reporter_->recordCall(genome.pegId(k), genome.pegFunction(k), roleNames_[r], roleColumn_[r], hits[s]);  // :147
            }
        }
        reporter_->closeGenome();                                         // :150
    }
    proteinsDone_ += batch.numPegs();
    residuesDone_ += batch.numResidues();
}

void ApplyKmerProcessor::runCommand() {
    GenomeDirectory genomes(inDir_);                                      // :116
    log_ << genomes.size() << " genomes found in input directory.\n";     // :117
    const std::vector<std::string>& files = genomes.files();
    // report column of every dense role id, once (ApplyKmerReporter.getRoleIdx, :92-95)
    roleColumn_.resize(roleNames_.size());
    for (size_t r = 0; r < roleNames_.size(); r++) roleColumn_[r] = reporter_->getRoleIdx(roleNames_[r]);
    // Ingest pipeline: two pinned batches alternate — the genomes of batch i+1 are parsed and packed by
    // `loadThreads_` threads while the GPU annotates batch i; reports are written in directory order (:118).
    PackedBatch bufs[2] = {PackedBatch(*engine_), PackedBatch(*engine_)};
    {
        // pinned buffers sized once from the file sizes (a file holds at least one byte per residue and ~40 per peg)
        uint64_t worst = 0;
        for (size_t b0 = 0; b0 < files.size(); b0 += (size_t)batchGenomes_) {
            uint64_t bytes = 0;
            for (size_t i = b0; i < std::min(files.size(), b0 + (size_t)batchGenomes_); i++) {
                struct stat st;
                if (stat(files[i].c_str(), &st) == 0) bytes += (uint64_t)st.st_size;
            }
            worst = std::max(worst, bytes);
        }
        const double tr = nowSeconds();
        if (worst < 0xf0000000ull)
            for (PackedBatch& b : bufs) b.reserve(worst, (size_t)(worst / 40 + 1024));
        log_ << "Two pinned ingest buffers for batches of up to " << worst / (1024 * 1024) << " MiB of genome files reserved in "
             << nowSeconds() - tr << " s.\n";
    }
    const double t0 = nowSeconds();
    const bool trace = getenv("KA_CLI_TRACE") != nullptr;
    double tLoad = 0, tFlush = 0, tWait = 0;
    auto loadBatch = [&](int which, size_t b0) {
        const size_t n = std::min(files.size() - b0, (size_t)batchGenomes_);
        const double t = nowSeconds();
        bufs[which].load(files.data() + b0, n, loadThreads_);
        tLoad += nowSeconds() - t;
        if (trace) log_ << "[ingest] batch at " << b0 << ": parse " << bufs[which].parseSeconds << " s, pack " << bufs[which].packSeconds << " s\n";
    };
    std::future<void> pending;
    if (!files.empty()) pending = std::async(std::launch::async, loadBatch, 0, (size_t)0);
    int cur = 0;
    for (size_t b0 = 0; b0 < files.size(); b0 += (size_t)batchGenomes_) {
        double t = nowSeconds();
        pending.get();
        tWait += nowSeconds() - t;
        const size_t nextStart = b0 + (size_t)batchGenomes_;
        if (nextStart < files.size()) pending = std::async(std::launch::async, loadBatch, cur ^ 1, nextStart);
        t = nowSeconds();
        flushBatch(bufs[cur]);
        tFlush += nowSeconds() - t;
        cur ^= 1;
    }
    if (trace) log_ << "[ingest] load (parse + pack, worker threads) " << tLoad << " s, main thread waited for it " << tWait
                    << " s, annotate + report " << tFlush << " s\n";
    reporter_->closeReport();                                             // :153
    reporter_->close();                                                   // :154
    const double dt = nowSeconds() - t0;
    log_ << "All done.  (t=" << nowSeconds() << " s)  " << proteinsDone_ << " proteins (" << residuesDone_ << " residues) of "
         << files.size() << " genomes, files to report in " << dt << " s = " << (dt > 0 ? proteinsDone_ / dt / 1e6 : 0.0)
         << " M proteins/s with " << loadThreads_ << " ingest threads.\n";
}

int ApplyKmerProcessor::run() {
    try {
        runCommand();
        return 0;
    } catch (const std::exception& e) {
        log_ << "EXECUTION ERROR: " << e.what() << "\n";
        return 1;
    }
}

}  // namespace theseed
