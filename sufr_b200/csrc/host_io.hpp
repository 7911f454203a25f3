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

// LCP of the pair (a, b) under the sort's own rules, on the transformed text (seam repair,
// sufr_builder.rs:893-902).  `mask` may be NULL, `q` = max_query_len or 0.
uint64_t host_pair_lcp(const uint8_t* text, uint64_t n, uint64_t a, uint64_t b, const SeedMaskInfo* mask, uint64_t q,
                       const uint64_t* n_ranges, uint64_t num_n_ranges);

// sufr_builder.rs:817-918
void write_sufr_file(const SufrB200Args& args, const SufrB200Result& r);

// util.rs:51-89
void read_sequence_file(const char* path, uint8_t delim, SufrB200Sequences* out);
void free_sequences(SufrB200Sequences* s);

}  // namespace sufr
