"""CPU: gt_gpu_gather takes the command line of the reference's gt_mpi_gather (tools/src/gt_mpi_gather.cc:438-545): the same
option spellings, the same refusal when -j is missing.  Nothing here reaches a CUDA call -- argument handling and the
configuration are read first."""
import json
import os
import subprocess

import pytest

from conftest import REPO


@pytest.fixture(scope="module")
def exe():
    import __graft_entry__ as g
    g.build_gpu_lib()
    return g.build_tools()


def test_version(exe):
    p = subprocess.run([exe, "--version"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.startswith("gdbgpu ")


def test_query_json_is_mandatory(exe):
    p = subprocess.run([exe, "--produce-Broad-GVCF"], capture_output=True, text=True)
    assert p.returncode != 0 and "Query JSON file (-j) is a mandatory argument - unspecified" in p.stderr


def test_unknown_option(exe):
    p = subprocess.run([exe, "--no-such-option"], capture_output=True, text=True)
    assert p.returncode != 0


@pytest.mark.parametrize("args", [["--json-config"], ["-j"], ["--query"]], ids=lambda a: a[0])
@pytest.mark.parametrize("command", ["--produce-Broad-GVCF", "--print-calls", "--print-csv", "--print-AC", "--produce-interesting-positions", "--produce-histogram", "--range-query"])
def test_reference_spellings_reach_the_configuration(exe, tmp_path, args, command):
    """every option of the reference tool is taken; the run then stops at the configuration, which the reference refuses too"""
    q = tmp_path / "q.json"
    q.write_text(json.dumps({"workspace": "/ws", "array": "a"}))
    p = subprocess.run([exe] + args + [str(q), command, "-O", "bu", "-w", "/other", "-A", "other", "-s", "1048576", "-r", "0", "-p", "1000", "--skip-query-on-root",
                                       "--output-format", "bu", "--workspace", "/w", "--array", "x", "--segment-size", "10", "--rank", "0", "--page-size", "10"],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "Must have one of \"query_column_ranges\"" in p.stderr, p.stderr


def test_other_output_formats_are_refused(exe, tmp_path):
    q = tmp_path / "q.json"
    q.write_text("{}")
    p = subprocess.run([exe, "-j", str(q), "--produce-Broad-GVCF", "-O", "z"], capture_output=True, text=True)
    assert p.returncode != 0 and "uncompressed BCF" in p.stderr
