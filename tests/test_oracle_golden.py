"""Pins the CPU oracle (oracle/oracle.cpp) to the unmodified reference: every file the compiled
reference binary wrote for 188 FASTA cases (146 of them the reference's own .tst cases, whose
stdout goldens the binary was checked against when the fixtures were generated) and the reads it
kept for 17 FASTQ cases must be reproduced exactly.  CPU only."""
import pytest

import golden_cmp
import oracle_lib as ol


def _ids(cases, kind):
    return [f"{c['fixture']}:{' '.join(c['args'])}" for c in cases if c["kind"] == kind]


def test_pattern_table_sizes():
    # SURVEY.md §8a row P [ran against the reference]: 38 default, 124 for the 4-variant set,
    # 44 for -c CCCTAAA, 308 for -x 2
    assert len(ol.expand_patterns(["CCCTAA", "TTAGGG"], 1, "CCCTAA")) == 38
    assert len(ol.expand_patterns(["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], 1, "CCCTAA")) == 124
    assert len(ol.expand_patterns(["CCCTAAA", "TTTAGGG"], 1, "CCCTAAA")) == 44
    assert len(ol.expand_patterns(["CCCTAA", "TTAGGG"], 2, "CCCTAA")) == 308
    t = dict(ol.expand_patterns(["CCCTAA", "TTAGGG"], 1, "CCCTAA"))
    assert t["CCCTAA"] is True and t["TTAGGG"] is False


def test_pattern_count_matches_reference_stderr(golden):
    cases, _ = golden
    for c in cases:
        cfg, _ = ol.config_from_args(c["args"], read_mode=c["kind"] == "fastq")
        n = len(ol.expand_patterns(cfg.raw_patterns(), cfg.edit_distance, cfg.canonical_fwd))
        assert f"Scanning {n} telomeric variants" in c["stderr"], (c["fixture"], c["args"])


def test_fasta_cases(golden):
    cases, fixtures = golden
    n = 0
    for c in cases:
        if c["kind"] != "fasta":
            continue
        cfg, flags = ol.config_from_args(c["args"])
        orc = ol.Oracle(cfg)
        golden_cmp.check_fasta_case(c, fixtures[c["fixture"]].tobytes(), orc.scan_segment, cfg, flags)
        n += 1
    assert n >= 180


def test_fastq_cases(golden):
    cases, fixtures = golden
    n = 0
    for c in cases:
        if c["kind"] != "fastq":
            continue
        cfg, _ = ol.config_from_args(c["args"], read_mode=True)
        orc = ol.Oracle(cfg)
        reads = ol.parse_fastq(fixtures[c["fixture"]].tobytes())
        kept = [name for name, seq in reads if orc.read_matches(seq)]
        assert kept == c["kept"], (c["fixture"], c["args"])
        n += 1
    assert n >= 15
