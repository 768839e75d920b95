// Genome.hpp — the slice of org.theseed.genome.{Genome,Feature,GenomeDirectory} and
// org.theseed.io.{TabbedLineReader,LineReader} (external SEEDtk classes, not in the reference
// repository) that the `apply` path touches: ApplyKmerProcessor.java:116-123 iterates
// `new GenomeDirectory(inDir)`, `genome.getPegs()`, `feat.getProteinTranslation()`; the
// reporters use `genome.getId()`, `feat.getId()`, `feat.getFunction()`
// (VerifyApplyKmerReporter.java:43-45).  Input files: GTO JSON (`*.gto`) or protein FASTA
// (`*.faa`, `*.fa`, `*.fasta`; genome id = file stem, comment = function).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace theseed {

struct FileNotFoundException : std::runtime_error { using std::runtime_error::runtime_error; };
struct ParseFailureException : std::runtime_error { using std::runtime_error::runtime_error; };
struct IOException : std::runtime_error { using std::runtime_error::runtime_error; };

class Feature {
public:
    Feature(std::string id, std::string type, std::string function, std::string protein)
        : id_(std::move(id)), type_(std::move(type)), function_(std::move(function)), protein_(std::move(protein)) {}
    const std::string& getId() const { return id_; }
    const std::string& getType() const { return type_; }
    const std::string& getFunction() const { return function_; }
    const std::string& getProteinTranslation() const { return protein_; }
    /** SEEDtk derives the feature type from the fid: fig|83333.1.peg.4 is a peg. */
    bool isPeg() const { return id_.find(".peg.") != std::string::npos; }
private:
    std::string id_, type_, function_, protein_;
};

class Genome {
public:
    /** Load a GTO (JSON) or a protein FASTA file. */
    explicit Genome(const std::string& path);
    const std::string& getId() const { return id_; }
    const std::string& getName() const { return name_; }
    /** Protein-encoding features, in file order. */
    std::vector<const Feature*> getPegs() const;
    const std::vector<Feature>& getFeatures() const { return features_; }
    std::string toString() const { return id_ + " (" + name_ + ")"; }
private:
    void loadGto(const std::string& text);
    void loadFasta(const std::string& text, const std::string& stem);
    std::string id_, name_;
    std::vector<Feature> features_;
    bool fromFasta_ = false;
};

/** Genome files of a directory, sorted by file name (GenomeDirectory keeps a sorted id set). */
class GenomeDirectory {
public:
    explicit GenomeDirectory(const std::string& dir);
    size_t size() const { return files_.size(); }
    const std::vector<std::string>& files() const { return files_; }
private:
    std::vector<std::string> files_;
};

/** Whole file into a string; throws FileNotFoundException. */
std::string readFile(const std::string& path);

/** Lines of a text file without their terminators (LineReader). */
std::vector<std::string> readLines(const std::string& path);

}  // namespace theseed
