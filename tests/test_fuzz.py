"""A few rounds of the differential fuzzers (tests/fuzz/fuzz_hostsim.py, tests/fuzz/fuzz_oracle_vs_reference.py, tests/fuzz/fuzz_cli_parser.py):
generated reads with substitutions, indels, N runs, low-complexity inserts, chimeras, lowercase and
random options.  The long runs are done by hand (DESIGN.md section 1c records them)."""
import os
import subprocess
import sys

import pytest

from oracle_binding import REF_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, rounds, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz", script), str(rounds), str(seed)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0 and b"ok:" in r.stdout, r.stdout.decode()[-2000:]


def test_product_stage_functions_against_oracle():
    _run("fuzz_hostsim.py", 8, 101)


def test_oracle_against_reference_binary():
    if not os.path.exists(os.path.join(REF_DIR, "centrifuger")):
        pytest.skip("oracle/_ref/centrifuger not built")
    _run("fuzz_oracle_vs_reference.py", 8, 202)


def test_cli_parser_against_reference_binary():
    if not os.path.exists(os.path.join(REF_DIR, "centrifuger")):
        pytest.skip("oracle/_ref/centrifuger not built")
    _run("fuzz_cli_parser.py", 25, 404)


def test_read_format_and_barcodes_against_reference_binary():
    if not os.path.exists(os.path.join(REF_DIR, "centrifuger")):
        pytest.skip("oracle/_ref/centrifuger not built")
    _run("fuzz_read_format.py", 20, 505)


def test_merge_readpair_cli_against_reference_binary():
    if not os.path.exists(os.path.join(REF_DIR, "centrifuger")):
        pytest.skip("oracle/_ref/centrifuger not built")
    _run("fuzz_merge_cli.py", 15, 606)


def test_block_parallel_ingest_against_serial_reader():
    """the CLI's block-parallel FASTQ reader (tiny byte ranges, 1 - 8 threads, irregular records mid-file) gives what
    its serial reader gives -- which test_cli_parser_against_reference_binary pins to the reference"""
    _run("fuzz_bulk_ingest.py", 40, 707)
