// selftest.cpp — host-side logic of the `apply` mirror without the GPU engine: genome
// readers and the two reporters, driven by files named on the command line.
//   kmers-anno-selftest <genome file> <roles.in.use> <VERIFY|APPLY> [calls.tsv]
// calls.tsv: peg index <TAB> role <TAB> hits (the calls to replay through recordFeature).
//   kmers-anno-selftest --json in.json out.json     parse and re-serialise a JSON document (JsonDoc.hpp)
//   kmers-anno-selftest --norm "function text"      print the normalised function (GeneCopyProcessor.hpp)
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "ApplyKmerReporter.hpp"
#include "Genome.hpp"
#include "GeneCopyProcessor.hpp"
#include "JsonDoc.hpp"

using namespace theseed;

int main(int argc, char** argv) {
    if (argc == 4 && std::string(argv[1]) == "--json") {
        try {
            std::ofstream(argv[3], std::ios::binary) << JsonValue::parse(readFile(argv[2])).dump();
            return 0;
        } catch (const std::exception& e) { std::cerr << "ERROR: " << e.what() << "\n"; return 1; }
    }
    if (argc == 3 && std::string(argv[1]) == "--norm") { std::cout << normalizeFunction(argv[2]) << "\n"; return 0; }
    if (argc < 4) { std::cerr << "usage: kmers-anno-selftest genome roles VERIFY|APPLY [calls.tsv]\n"; return 2; }
    try {
        Genome genome(argv[1]);
        auto pegs = genome.getPegs();
        std::cerr << "genome " << genome.getId() << " features " << genome.getFeatures().size() << " pegs " << pegs.size() << "\n";
        auto reporter = ApplyKmerReporter::create(ApplyKmerReporter::parseType(argv[3]), std::cout);
        reporter->initReport(argv[2]);
        std::map<size_t, std::pair<std::string, int>> calls;
        if (argc > 4)
            for (const std::string& line : readLines(argv[4])) {
                std::istringstream ss(line);
                size_t idx; std::string role; int hits;
                if (ss >> idx >> role >> hits) calls[idx] = {role, hits};
            }
        reporter->openGenome(genome);
        for (size_t i = 0; i < pegs.size(); i++) {
            auto it = calls.find(i);
            if (it != calls.end()) reporter->recordFeature(*pegs[i], it->second.first, it->second.second);
        }
        reporter->closeGenome();
        reporter->closeReport();
        reporter->close();
        // echo what the reader saw, for the test to compare
        std::ofstream dump(std::string(argv[1]) + ".dump");
        for (const Feature* f : pegs) dump << f->getId() << "\t" << f->getFunction() << "\t" << f->getProteinTranslation() << "\n";
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
