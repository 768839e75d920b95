// App.cpp — command dispatch, mirror of proteins/kmers/anno/App.java:51-84 for the verbs of the
// GPU hot path.  `build`, `apply` and `genes` run on the engine; the reference's other verbs are outside the
// scope of this engine and are reported as such.
#include <iostream>
#include <cstdlib>
#include <string>
#include <vector>

#include "ApplyKmerProcessor.hpp"
#include "BuildKmerProcessor.hpp"
#include "GeneCopyProcessor.hpp"

using namespace theseed;

static const char* kCommands[][2] = {
    {"build", "build a discriminating-kmer database from annotated genomes (GPU engine)"},
    {"apply", "apply a discriminating-kmer database to genomes (GPU engine)"},
    {"genes", "copy aliases between close genomes using protein kmer distance (GPU engine)"},
};

static void showCommands() {
    std::cerr << "Available commands:\n";
    for (auto& c : kCommands) std::cerr << "  " << c[0] << "\t" << c[1] << "\n";
}

int main(int argc, char** argv) {
    if (argc < 2) { showCommands(); return 1; }
    // stdout carries the report (ApplyKmerProcessor.java:94): NCCL's start-up lines (--table-mode 2) go to stderr.
    // Done here, in the still single-threaded host, not inside the library.
    setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    setenv("KA_NCCL_STDOUT_TO_STDERR", "1", 0);
    std::string command = argv[1];
    std::vector<std::string> newArgs(argv + 2, argv + argc);   // App.java:54-55
    if (command == "apply") {
        ApplyKmerProcessor processor;                          // App.java:62
        if (!processor.parseCommand(newArgs)) return 1;        // App.java:81
        return processor.run();                                // App.java:82
    }
    if (command == "build") {
        BuildKmerProcessor processor;                          // App.java:61
        if (!processor.parseCommand(newArgs)) return 1;
        return processor.run();
    }
    if (command == "genes") {
        GeneCopyProcessor processor;                           // App.java:71
        if (!processor.parseCommand(newArgs)) return 1;
        return processor.run();
    }
    if (command == "-h" || command == "--help") { showCommands(); return 0; }
    std::cerr << "Invalid command " << command << ".\n";       // App.java:76
    showCommands();
    return 1;
}
