// VCF / BCF header lines as the reference's host side keeps them (htslib bcf_hdr_t), without htslib: the template header of
// the configuration is parsed, the queried fields / the contigs of the vid mapping / the samples of the query are added, and
// the text comes out byte for byte as bcf_hdr_format() writes it -- so that the header at the head of a BCF stream produced
// here equals the one the unmodified reference writes (tests/golden/bcf/*.bcf).  Restated behaviour (htslib of GenomicsDB
// v0.10.2, dependencies/htslib/vcf.c): bcf_hdr_parse_line :289-391, bcf_hdr_register_hrec :407-545 (dictionary indices, IDX),
// bcf_hdr_add_hrec :547-588 (duplicates), bcf_hdr_parse_required_sample_line :645-718 (PASS first), bcf_hdr_remove :736-795,
// _bcf_hrec_format / bcf_hdr_format :1837-1881, bcf_hdr_add_sample :89-118; indexed FASTA: faidx.c (fai_load,
// faidx_fetch_seq) and bgzf.c (.gzi random access).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace genomicsdb_b200 {

enum { HL_FLT = 0, HL_INFO = 1, HL_FMT = 2, HL_CTG = 3, HL_STR = 4, HL_GEN = 5 };   // BCF_HL_*
enum { HT_FLAG = 0, HT_INT = 1, HT_REAL = 2, HT_STR = 3 };                          // BCF_HT_*
enum { VL_FIXED = 0, VL_VAR = 1, VL_A = 2, VL_G = 3, VL_R = 4 };                    // BCF_VL_*

struct VCFHeaderRec {
  int type = HL_GEN;
  std::string key;                                             // "INFO", "contig", "fileformat", ...
  std::string value;                                           // generic lines (##key=value)
  bool structured = false;                                     // ##key=<k=v,...>
  std::vector<std::pair<std::string, std::string>> kv;         // values keep their quotes
  int find_key(const std::string& k) const { for (size_t i = 0; i < kv.size(); ++i) if (kv[i].first == k) return (int)i; return -1; }
};

class VCFHeaderLines {
 public:
  // write_mode: what bcf_hdr_init("w") starts with (##fileformat=VCFv4.2 and the PASS filter)
  explicit VCFHeaderLines(bool write_mode = false);
  // a whole header text (bcf_hdr_parse): false when the sample line is required and missing
  bool parse(const std::string& text, bool sample_line_required = false);
  bool append(const std::string& line);                        // bcf_hdr_append
  void remove(int type, const std::string& id);                // bcf_hdr_remove(hdr, type, key)
  int id2int(const std::string& id) const;                     // bcf_hdr_id2int(hdr, BCF_DT_ID, id), -1: unknown
  bool exists(int type, const std::string& id) const;          // idinfo exists and its line is still there
  const VCFHeaderRec* hrec(int type, const std::string& id) const;
  int id2type(int type, const std::string& id) const;          // BCF_HT_*
  int id2length(int type, const std::string& id) const;        // BCF_VL_*
  int id2number(int type, const std::string& id) const;
  int name2id(const std::string& contig) const;                // bcf_hdr_name2id, -1: unknown
  bool add_sample(const std::string& name);                    // false: duplicate
  size_t nsamples() const { return samples_.size(); }
  const std::vector<std::string>& samples() const { return samples_; }
  std::string format(bool is_bcf) const;                       // bcf_hdr_format: IDX= only for BCF
  // dictionaries as (name, index) in index order
  std::vector<std::pair<std::string, int>> ids() const;
  std::vector<std::pair<std::string, int>> contigs() const;
 private:
  struct IdInfo { int id = -1; uint32_t info[3] = {15, 15, 15}; VCFHeaderRec* rec[3] = {nullptr, nullptr, nullptr}; };
  struct CtgInfo { int id = -1; VCFHeaderRec* rec = nullptr; };
  static std::unique_ptr<VCFHeaderRec> parse_line(const char* line, int* len);
  bool register_rec(VCFHeaderRec* r);                          // bcf_hdr_register_hrec
  int add_rec(std::unique_ptr<VCFHeaderRec> r);                // bcf_hdr_add_hrec
  std::vector<std::unique_ptr<VCFHeaderRec>> recs_;
  std::map<std::string, IdInfo> ids_;
  std::map<std::string, CtgInfo> ctgs_;
  std::vector<bool> id_used_, ctg_used_;                       // dictionary slots taken (conflicting IDX detection)
  int n_ids_ = 0, n_ctgs_ = 0;
  std::vector<std::string> samples_;
};

// An indexed FASTA (<file>.fai; bgzip-compressed files also need <file>.gzi), read the way the reference reads it
// (ReferenceGenomeInfo, vcf_adapter.cc:31-58: one base at a time through a 32 KB window).
class IndexedFasta {
 public:
  IndexedFasta() = default;
  ~IndexedFasta();
  IndexedFasta(const IndexedFasta&) = delete;
  IndexedFasta& operator=(const IndexedFasta&) = delete;
  void open(const std::string& path);                          // throws std::runtime_error
  bool is_open() const { return fd_ >= 0; }
  // base at 0-based position `pos` of `contig`; 'N' past the end, throws for an unknown contig
  char base_at(const std::string& contig, int64_t pos);
 private:
  struct Entry { int64_t length, offset, line_bases, line_width; };
  void read_raw(int64_t uoffset, char* dst, size_t n);          // n bytes of the (uncompressed) file at uoffset
  std::map<std::string, Entry> index_;
  int fd_ = -1;
  bool bgzf_ = false;
  std::vector<std::pair<uint64_t, uint64_t>> gzi_;              // (compressed, uncompressed) offset of every block
  std::vector<uint8_t> block_; int64_t block_uoff_ = -1;        // the inflated block last touched
  std::string win_contig_; int64_t win_pos_ = -1; std::string win_;
};

}  // namespace genomicsdb_b200
