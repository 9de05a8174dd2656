/*
 * centrifuger_b200_build.h -- C ABI of the GPU index builder.
 *
 * Replaces, for nucleotide collections, what `centrifuger-build` does between reading the genomes
 * and writing <prefix>.1.cfr:
 *     FMBuilder::Build                    compactds/FMBuilder.hpp:444
 *     Builder::TransformSampledSAToSeqId  Builder.hpp:27
 *     FMIndex::Init / Save                compactds/FMIndex.hpp:256 / :571
 *     Sequence_RunBlock::Init             compactds/Sequence_RunBlock.hpp:231
 * The file written is byte-identical to the reference builder's for the same text and options
 * (offrate / ftabchars / rbbwt-b).  Plain C, plain pointers and sizes.  Every function returns 0 or a
 * negative cfr_status (centrifuger_b200.h); cfr_build_last_error() gives the message.  There is no CPU
 * path: without a CUDA device the builder fails with CFR_ERR_CUDA.
 */
#ifndef CENTRIFUGER_B200_BUILD_H
#define CENTRIFUGER_B200_BUILD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* replaces the _FMBuilderParam fields centrifuger-build exposes (CentrifugerBuild.cpp:20-26,80-81) */
typedef struct {
  int32_t sample_rate;      /* 2^--offrate        [16] */
  int32_t precompute_width; /* --ftabchars        [10] */
  uint64_t rbbwt_b;         /* --rbbwt-b, 0 = automatic, 1 = no run blocks [0] */
  uint64_t max_batch_rows;  /* suffixes sorted per batch, 0 = sized from the free HBM */
  int32_t device;
  int32_t verbose;          /* progress lines on stderr */
} cfr_build_params;

typedef struct {
  uint64_t batches;
  uint64_t max_tie_depth_bases; /* longest common prefix the sorter had to look through (rounded up to 31) */
  double sort_seconds, derive_seconds, runblock_seconds, write_seconds;
} cfr_build_stats;

void cfr_build_default_params(cfr_build_params *p);
const char *cfr_build_last_error(void);

/* codes: the concatenated genomes, one byte per base (0..3 = A,C,G,T), n bytes, HOST memory, exactly what
 * SequenceCompactor::Compact leaves of the FASTA records (non-ACGT characters dropped, SequenceCompactor.hpp:59-84).
 * genome_lens / genome_seq_ids: length and sequence id of each of the n_genomes records, in text order
 * (Builder::Build's genomeLens / genomeSeqIds, Builder.hpp:118-163).  Writes out_path (the .1.cfr file). */
int cfr_build_fm_index(const uint8_t *codes, uint64_t n, const uint64_t *genome_lens, const uint64_t *genome_seq_ids,
                       uint64_t n_genomes, const cfr_build_params *p, const char *out_path, cfr_build_stats *stats);

/* The same for a synthetic collection generated on the device: species x strains genomes of genome_len
 * bases; a species is uniform random, a strain differs from it in div_ppm of a million positions.  Genome g
 * (= species * strains + strain) has sequence id g.  Used for the 20 / 140 Gbp bench workloads, whose text
 * never exists on the host. */
int cfr_build_synthetic_fm_index(uint64_t species, uint64_t strains, uint64_t genome_len, uint64_t div_ppm, uint64_t seed,
                                 const cfr_build_params *p, const char *out_path, cfr_build_stats *stats);

/* bases [offset, offset + count) of one synthetic genome, as codes 0..3 (host; for drawing reads) */
void cfr_synth_bases(uint64_t species_index, uint64_t strain_index, uint64_t offset, uint64_t count, uint64_t div_ppm,
                     uint64_t seed, uint8_t *out_codes);

/* n fragments of `length` bases each: fragment i = bases [offset[i], offset[i] + length) of genome genome[i]
 * (= species * strains + strain); out_codes holds n rows of `length` codes.  Host threads. */
void cfr_synth_fragments(const uint64_t *genome, const uint64_t *offset, uint64_t n, uint64_t length, uint64_t strains,
                         uint64_t div_ppm, uint64_t seed, uint8_t *out_codes);

#ifdef __cplusplus
}
#endif
#endif
