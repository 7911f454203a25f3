// `sufr-b200 create`: command-line mirror of `sufr create` (reference: sufr/src/lib.rs:83-125 flag
// definitions, :321-371 create(), sufr/src/main.rs:8-41 global flags and error exit).
// Only the `create` subcommand (alias `cr`) exists here; the query commands stay with the reference.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/sufr_b200.h"

namespace {

struct CreateArgs {  // sufr/src/lib.rs:85-125
    std::string input;
    uint64_t num_partitions = 16;
    bool has_mql = false;
    uint64_t max_query_len = 0;
    std::string output;
    bool has_output = false;
    bool is_dna = false, allow_ambiguity = false, ignore_softmask = false;
    char sequence_delimiter = '%';
    std::string seed_mask;
    bool has_mask = false;
    uint64_t random_seed = 42;
    // global flags (sufr/src/lib.rs:35-45); threads is accepted and unused on the GPU path
    int threads = 0;
    std::string log_level;
    std::string log_file;
    std::vector<int> devices{0};
    uint32_t index_bits = 0;  // 0 = the reference's dispatch (u32 below u32::MAX bytes, suffix_array.rs:460-470)
};

[[noreturn]] void usage_error(const std::string& msg) {
    fprintf(stderr, "error: %s\n\nUsage: sufr-b200 create [OPTIONS] <INPUT>\n\nFor more information, try '--help'.\n",
            msg.c_str());
    exit(2);  // clap's usage-error exit code
}

void print_help() {
    puts("Create sufr file\n\n"
         "Usage: sufr-b200 [-t THREADS] [-l LOG] [--log-file FILE] create [OPTIONS] <INPUT>\n\n"
         "Arguments:\n  <INPUT>  Input file\n\nOptions:\n"
         "  -n, --num-partitions <NUM_PARTS>   Subproblem count [default: 16]\n"
         "  -m, --max-query-len <CONTEXT>      Max context\n"
         "  -o, --output <OUTPUT>              Output file\n"
         "  -d, --dna                          Input is DNA\n"
         "  -a, --allow-ambiguity              Allow suffixes starting with ambiguity codes\n"
         "  -i, --ignore-softmask              Ignore suffixes in soft-mask/lowercase regions\n"
         "  -D, --sequence-delimiter <DELIM>   Character to separate sequences [default: %]\n"
         "  -s, --seed-mask <MASK>             Spaced seeds mask\n"
         "  -r, --random-seed <RANDSEED>       Random seed [default: 42]\n"
         "      --device <ORDINAL>             CUDA device [default: 0]\n"
         "      --devices <LIST>               CUDA devices that share the build, e.g. 0-7 or 0,2,5 (one key-range shard each)\n"
         "      --index-bits <32|64>           Width of the SA / LCP entries [default: as the reference: 32 below 4 Gi bytes]\n"
         "  -h, --help                         Print help");
}

uint64_t parse_u64(const std::string& flag, const char* v) {
    char* end = nullptr;
    if (!v || !*v || *v == '-') usage_error("invalid value '" + std::string(v ? v : "") + "' for '" + flag + "'");
    unsigned long long x = strtoull(v, &end, 10);
    if (*end) usage_error("invalid value '" + std::string(v) + "' for '" + flag + "'");
    return x;
}

std::string file_stem(const std::string& path) {  // PathBuf::file_stem (sufr/src/lib.rs:334-340)
    size_t slash = path.find_last_of('/');
    std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
    if (base.empty()) return "out";
    size_t dot = base.find_last_of('.');
    if (dot == std::string::npos || dot == 0) return base;
    return base.substr(0, dot);
}

std::vector<int> parse_devices(const std::string& flag, const char* v) {  // "0-7", "0,2,5", "3"
    std::vector<int> out;
    std::string s = v ? v : "";
    size_t i = 0;
    while (i <= s.size()) {
        size_t j = s.find(',', i);
        if (j == std::string::npos) j = s.size();
        std::string part = s.substr(i, j - i);
        size_t dash = part.find('-');
        if (part.empty()) usage_error("invalid value '" + s + "' for '" + flag + "'");
        if (dash == std::string::npos) {
            out.push_back((int)parse_u64(flag, part.c_str()));
        } else {
            int lo = (int)parse_u64(flag, part.substr(0, dash).c_str()), hi = (int)parse_u64(flag, part.substr(dash + 1).c_str());
            if (hi < lo || hi - lo > 63) usage_error("invalid value '" + s + "' for '" + flag + "'");
            for (int d = lo; d <= hi; d++) out.push_back(d);
        }
        i = j + 1;
    }
    return out;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char** argv) {
    CreateArgs a;
    std::vector<std::string> pos;
    bool saw_create = false;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        auto value = [&](const std::string& flag) -> const char* {
            size_t eq = s.find('=');
            if (s.rfind("--", 0) == 0 && eq != std::string::npos) return argv[i] + eq + 1;
            if (i + 1 >= argc) usage_error("a value is required for '" + flag + "' but none was supplied");
            return argv[++i];
        };
        std::string key = s.rfind("--", 0) == 0 ? s.substr(0, s.find('=')) : s;
        if (key == "-h" || key == "--help") { print_help(); return 0; }
        else if (key == "-t" || key == "--threads") a.threads = (int)parse_u64(key, value(key));
        else if (key == "-l" || key == "--log") a.log_level = value(key);
        else if (key == "--log-file") a.log_file = value(key);
        else if (key == "-n" || key == "--num-partitions") a.num_partitions = parse_u64(key, value(key));
        else if (key == "-m" || key == "--max-query-len") { a.max_query_len = parse_u64(key, value(key)); a.has_mql = true; }
        else if (key == "-o" || key == "--output") { a.output = value(key); a.has_output = true; }
        else if (key == "-d" || key == "--dna") a.is_dna = true;
        else if (key == "-a" || key == "--allow-ambiguity") a.allow_ambiguity = true;
        else if (key == "-i" || key == "--ignore-softmask") a.ignore_softmask = true;
        else if (key == "-D" || key == "--sequence-delimiter") {
            const char* v = value(key);
            if (strlen(v) != 1) usage_error("invalid value '" + std::string(v) + "' for '--sequence-delimiter <DELIM>'");
            a.sequence_delimiter = v[0];
        }
        else if (key == "-s" || key == "--seed-mask") { a.seed_mask = value(key); a.has_mask = true; }
        else if (key == "-r" || key == "--random-seed") a.random_seed = parse_u64(key, value(key));
        else if (key == "--device") a.devices = {(int)parse_u64(key, value(key))};
        else if (key == "--devices") a.devices = parse_devices(key, value(key));
        else if (key == "--index-bits") {
            a.index_bits = (uint32_t)parse_u64(key, value(key));
            if (a.index_bits != 32 && a.index_bits != 64) usage_error("invalid value for '--index-bits <32|64>'");
        }
        else if (!s.empty() && s[0] == '-' && s.size() > 1) usage_error("unexpected argument '" + s + "' found");
        else if (!saw_create && (s == "create" || s == "cr")) saw_create = true;
        else pos.push_back(s);
    }
    if (!saw_create) {
        if (argc <= 1) { print_help(); return 2; }
        usage_error("unrecognized subcommand (only 'create' is provided by sufr-b200)");
    }
    if (pos.size() != 1) usage_error(pos.empty() ? "the following required arguments were not provided:\n  <INPUT>"
                                                 : "unexpected argument '" + pos[1] + "' found");
    a.input = pos[0];
    if (a.has_mql && a.has_mask)  // clap conflicts_with (sufr/src/lib.rs:95)
        usage_error("the argument '--max-query-len <CONTEXT>' cannot be used with '--seed-mask <MASK>'");

    const bool info = a.log_level == "info" || a.log_level == "debug";
    FILE* logf = stdout;
    if (info && !a.log_file.empty()) {
        logf = fopen(a.log_file.c_str(), "w");
        if (!logf) { fprintf(stderr, "Error: %s: cannot open log file\n", a.log_file.c_str()); return 1; }
    }

    double t0 = now_s();
    SufrB200Sequences seqs;
    if (sufr_b200_read_sequence_file(a.input.c_str(), (uint8_t)a.sequence_delimiter, &seqs) != SUFR_B200_OK) {
        fprintf(stderr, "Error: %s\n", sufr_b200_last_error());
        return 1;
    }
    if (info) fprintf(logf, "Read input of len %llu in %.3fs\n", (unsigned long long)seqs.seq_len, now_s() - t0);

    std::string outfile = a.has_output ? a.output : file_stem(a.input) + ".sufr";
    SufrB200Args args;
    memset(&args, 0, sizeof(args));
    args.text = seqs.seq;
    args.text_len = seqs.seq_len;
    args.path = outfile.c_str();
    args.low_memory = 1;
    args.has_max_query_len = a.has_mql;
    args.max_query_len = a.max_query_len;
    args.is_dna = a.is_dna;
    args.allow_ambiguity = a.allow_ambiguity;
    args.ignore_softmask = a.ignore_softmask;
    args.sequence_starts = seqs.start_positions;
    args.sequence_names = (const char* const*)seqs.sequence_names;
    args.num_sequences = seqs.num_sequences;
    args.num_partitions = a.num_partitions;
    args.seed_mask = a.has_mask ? a.seed_mask.c_str() : nullptr;
    args.random_seed = a.random_seed;
    args.rank = 0;
    args.world_size = 1;

    t0 = now_s();
    SufrB200Result res;
    int rc = (a.devices.size() == 1 && a.index_bits == 0)
                 ? sufr_b200_create(&args, a.devices[0], &res)
                 : sufr_b200_create_multi(&args, a.devices.data(), (int)a.devices.size(), a.index_bits, &res);
    if (rc != SUFR_B200_OK) {
        fprintf(stderr, "Error: %s\n", sufr_b200_last_error());  // sufr/src/main.rs:9-12
        sufr_b200_sequences_free(&seqs);
        return 1;
    }
    if (info) {
        const SufrB200Timings& t = res.timings;
        fprintf(logf, "Encoded text (alphabet %u, %u bits/symbol) in %.3fms\n", res.alphabet_size, res.bits_per_symbol,
                t.encode_ms);
        fprintf(logf, "Sorted %llu suffixes on %zu GPU%s in %.3fms (keys %.3f, sort %.3f, refine %.3f [%u word / %u doubling rounds], "
                      "lcp %.3f, finish %.3f; h2d %.3f, d2h %.3f; %llu kernel launches)\n",
                (unsigned long long)res.num_suffixes, a.devices.size(), a.devices.size() == 1 ? "" : "s", t.total_ms, t.keys_ms, t.sort_ms, t.refine_ms, res.refine_rounds,
                res.doubling_rounds, t.lcp_ms, t.finish_ms, t.h2d_ms, t.d2h_ms, (unsigned long long)res.kernel_launches);
        struct stat sb;
        unsigned long long bytes = stat(outfile.c_str(), &sb) == 0 ? (unsigned long long)sb.st_size : 0;
        fprintf(logf, "Wrote %llu byte%s to '%s' in %.3fs\n", bytes, bytes == 1 ? "" : "s", outfile.c_str(), now_s() - t0);
    }
    sufr_b200_result_free(nullptr, &res);
    sufr_b200_sequences_free(&seqs);
    if (logf != stdout) fclose(logf);
    return 0;
}
