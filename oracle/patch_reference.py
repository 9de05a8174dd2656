#!/usr/bin/env python
"""Test infrastructure ONLY: proves INTEGRATION.md's "Option 2" -- the ~60-line patch a maintainer of the reference would
apply to put libcfrb200.so behind the reference's own batch loop -- by really applying it.

    python oracle/patch_reference.py [REF_DIR] [OUT_BINARY]

reads REF_DIR/CentrifugerClass.cpp where it lies (default /root/reference), applies three textual edits IN MEMORY,
writes the patched translation unit to a temporary directory outside the repo, and compiles it with g++ against the
unmodified headers of REF_DIR and include/centrifuger_b200.h into OUT_BINARY (default oracle/_ref/centrifuger_patched,
git-ignored, shipped to the GPU box).  No reference source is copied into the repo.  The result is the REFERENCE's binary --
its argv handling, ReadFiles / kseq ingest, ResultWriter -- with Classifier::Query replaced by cfr_submit_batch on the GPU;
tests/test_gpu_parity.py::test_patched_reference_binary compares its TSV with the unmodified binary's goldens.

The edits (line numbers of the reference, CentrifugerClass.cpp):
  A  after the #include block (:19): #include "centrifuger_b200.h" and the helper ClassifyBatchOnGpu -- the text of
     INTEGRATION.md, Option 2
  B  after classifier.Init / IsProteinDatabase (:595-596): open the GPU handle from the same _classifierParam
  C  the per-batch fan-out of the single-reader branch (:681-688): one call instead of pthread_create / pthread_join
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

HELPER = r'''
#include "centrifuger_b200.h"            // link with -lcfrb200

// One batch through the GPU instead of ClassifyReads_Thread (CentrifugerClass.cpp:238-340): pack the _Read arrays into the
// SoA batch, classify (DUST masking included, as in ClassifyReads_Thread), convert the records back to _classifierResult
// with the library's Taxonomy look-ups (the Classifier keeps its own Taxonomy private).
static void ClassifyBatchOnGpu(cfr_handle *gpu, struct _Read *readBatch, struct _Read *readBatch2, int batchSize,
                               struct _classifierResult *results, bool expanded)
{
  const int k = (int)cfr_index_info(gpu, 24) ;  // id slots per read: -k, or the cap of -k 0
  std::string s1, s2 ; std::vector<uint64_t> o1(1, 0), o2(1, 0) ;
  for (int i = 0 ; i < batchSize ; ++i)
  {
    s1 += readBatch[i].seq ;  o1.push_back(s1.size()) ;
    if (readBatch2) { s2 += readBatch2[i].seq ; o2.push_back(s2.size()) ; }
  }
  cfr_read_batch b = { (uint64_t)batchSize, s1.data(), o1.data(),
                       readBatch2 ? s2.data() : NULL, readBatch2 ? o2.data() : NULL } ;
  std::vector<cfr_result> r(batchSize) ; std::vector<uint64_t> ids((size_t)batchSize * k) ;
  int ticket = -1 ;
  if (cfr_submit_batch(gpu, &b, r.data(), ids.data(), NULL, &ticket) != CFR_OK || cfr_wait_batch(gpu, ticket) != CFR_OK)
  {
    Utils::PrintLog("ERROR: %s", cfr_last_error()) ; exit(EXIT_FAILURE) ;
  }
  std::vector<uint32_t> expCnt ; std::vector<uint64_t> expOff, expIds ;
  if (expanded)  // --expand-taxid: the ids promoted into each reported id (Classifier.hpp:807-838)
  {
    expCnt.resize((size_t)batchSize * k) ; expOff.resize(batchSize) ; expIds.resize(1024) ;
    uint64_t need = 0 ;
    int st = cfr_fetch_expanded(gpu, ticket, expCnt.data(), expOff.data(), expIds.data(), expIds.size(), &need) ;
    if (st != CFR_OK && need > expIds.size())
    {
      expIds.resize(need) ;
      st = cfr_fetch_expanded(gpu, ticket, expCnt.data(), expOff.data(), expIds.data(), expIds.size(), &need) ;
    }
    if (st != CFR_OK) { Utils::PrintLog("ERROR: %s", cfr_last_error()) ; exit(EXIT_FAILURE) ; }
  }
  for (int i = 0 ; i < batchSize ; ++i)            // -> _classifierResult (Classifier.hpp:41-59)
  {
    results[i].Clear() ;
    results[i].score = r[i].score ;  results[i].secondaryScore = r[i].secondary_score ;
    results[i].hitLength = r[i].hit_length ;  results[i].queryLength = r[i].query_length ;
    uint64_t at = expanded ? expOff[i] : 0 ;
    for (int j = 0 ; j < r[i].n_assign && j < k ; ++j)
    {
      const uint64_t id = ids[(size_t)i * k + j] ;
      if (r[i].by_rank)                            // Classifier.hpp:818-820
      {
        results[i].seqStrNames.push_back(cfr_rank_name(gpu, id)) ;
        results[i].taxIds.push_back(cfr_orig_taxid(gpu, id)) ;
      }
      else                                         // Classifier.hpp:790-791
      {
        results[i].seqStrNames.push_back(cfr_seq_name(gpu, id)) ;
        results[i].taxIds.push_back(cfr_orig_taxid(gpu, cfr_seq_taxid(gpu, id))) ;
      }
      if (expanded)                                // Classifier.hpp:794, :821-838: original ids joined by commas
      {
        std::string e ;
        const uint32_t cnt = expCnt[(size_t)i * k + j] ;
        for (uint32_t q = 0 ; q < cnt ; ++q)
        {
          if (q) e += "," ;
          e += std::to_string((unsigned long)cfr_orig_taxid(gpu, expIds[at + q])) ;
        }
        at += cnt ;
        results[i].expandedTaxIdStrings.push_back(e) ;
      }
    }
  }
}
'''

OPEN = r'''
  cfr_handle *gpu = NULL ;
  if (!protein && !mergeReadPair)  // (a protein index and --merge-readpair keep the CPU path)
  {
    cfr_params gp ; cfr_default_params(&gp) ;
    gp.max_result                  = classifierParam.maxResult ;             // -k
    gp.min_hit_len                 = classifierParam.minHitLen ;             // --min-hitlen (0 = infer)
    gp.max_result_per_hit_factor   = classifierParam.maxResultPerHitFactor ; // --hitk-factor
    gp.consider_secondary_hit_len  = classifierParam.considerSecondaryHitLen ;
    gp.consider_secondary_score_factor = classifierParam.considerSecondaryScoreFactor ;
    gp.dust                        = dust ? 1 : 0 ;                          // --no-dust
    gp.expand_taxid                = classifierParam.outputExpandedResult ? 1 : 0 ;
    gp.max_batch_reads             = 1 << 20 ;
    if (cfr_open(idxPrefix, &gp, /*device=*/0, &gpu) != CFR_OK)
    {
      Utils::PrintLog("ERROR: %s", cfr_last_error()) ;
      exit(EXIT_FAILURE) ;                                                   // same failure style as :121-125
    }
  }
'''

FANOUT_OLD = '''      for ( i = 0 ; i < classificationThreadCnt ; ++i )
      {
        args[i].batchSize = batchSize ;
        pthread_create( &threads[i], &attr, ClassifyReads_Thread, (void *)&args[i] ) ;
      }

      for ( i = 0 ; i < classificationThreadCnt ; ++i )
        pthread_join( threads[i], NULL ) ;

      for (i = 0 ; i < batchSize ; ++i)
        resWriter.Output(readBatch[i].id, readBatch[i].seq, readBatch[i].qual,'''

FANOUT_NEW = '''      if (gpu)
        ClassifyBatchOnGpu(gpu, readBatch, hasMate ? readBatch2 : NULL, batchSize, classifierBatchResults,
            classifierParam.outputExpandedResult) ;
      else
      {
      for ( i = 0 ; i < classificationThreadCnt ; ++i )
      {
        args[i].batchSize = batchSize ;
        pthread_create( &threads[i], &attr, ClassifyReads_Thread, (void *)&args[i] ) ;
      }

      for ( i = 0 ; i < classificationThreadCnt ; ++i )
        pthread_join( threads[i], NULL ) ;
      }

      for (i = 0 ; i < batchSize ; ++i)
        resWriter.Output(readBatch[i].id, readBatch[i].seq, readBatch[i].qual,'''


def replace_once(text, old, new, what):
    if text.count(old) != 1:
        raise SystemExit("patch_reference: anchor for %s found %d times (the reference changed?)" % (what, text.count(old)))
    return text.replace(old, new)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "_ref", "centrifuger_patched")
    src = open(os.path.join(ref, "CentrifugerClass.cpp")).read()
    src = replace_once(src, '#include "Dustmasker.hpp"\n', '#include "Dustmasker.hpp"\n' + HELPER, "edit A (helper)")
    anchor_b = "  protein = classifier.IsProteinDatabase() ;\n"
    src = replace_once(src, anchor_b, anchor_b + OPEN, "edit B (open)")
    src = replace_once(src, FANOUT_OLD, FANOUT_NEW, "edit C (fan-out)")
    libdir = os.path.join(ROOT, "centrifuger_b200")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="cfr_patched_ref_") as d:
        tu = os.path.join(d, "CentrifugerClass_b200.cpp")
        with open(tu, "w") as f:
            f.write(src)
        cmd = ["g++", "-w", "-O3", "-msse4.2", "-I" + ref, "-I" + os.path.join(ROOT, "include"), "-o", out, tu,
               "-L" + libdir, "-lcfrb200", "-Wl,-rpath," + libdir, "-Wl,-rpath,$ORIGIN/../../centrifuger_b200", "-lpthread", "-lz"]
        subprocess.check_call(cmd)
    print(out)


if __name__ == "__main__":
    main()
