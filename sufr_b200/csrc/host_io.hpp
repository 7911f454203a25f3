// Host-side pieces of the create path: SeedMask, the `.sufr` writer and FASTA/FASTQ ingest.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sufr_b200.h"

namespace sufr {

// libsufr/src/types.rs:36-200
struct SeedMaskInfo {
    std::string mask;
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> positions;
    uint64_t weight = 0;
};
bool parse_seed_mask(const char* mask, SeedMaskInfo& out);
uint64_t lcp_full_offset(uint64_t lcp, const SeedMaskInfo& m);  // util.rs:19-37

// sufr_builder.rs:817-918
void write_sufr_file(const SufrB200Args& args, const SufrB200Result& r);

// The parts of a version-6 `.sufr` file around the three big sections (sufr_builder.rs:826-867, :909).
struct SufrFrame {
    std::vector<uint8_t> head;  // everything before the text
    std::vector<uint8_t> tail;  // bincode Vec<String> of the sequence names
    uint64_t text_pos = 0, sa_pos = 0, lcp_pos = 0, names_pos = 0;
};
SufrFrame make_sufr_frame(const SufrB200Args& args, uint32_t index_bits, uint64_t text_len, uint64_t total_suffixes);
void pwrite_all(int fd, const void* buf, size_t len, uint64_t off, const std::string& path);

// util.rs:51-89
void read_sequence_file(const char* path, uint8_t delim, SufrB200Sequences* out);
void free_sequences(SufrB200Sequences* s);

}  // namespace sufr
