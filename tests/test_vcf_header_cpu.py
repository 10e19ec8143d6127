"""CPU: VCFHeaderLines -- htslib's header bookkeeping restated in the library (bcf_hdr_parse / append / add_sample / sync / format,
the dictionaries of IDs and contigs) -- against the reference's own htslib (oracle/_ref/bin/gdb_ref_dump --mode header) on header
texts a template may hold: IDX fields, repeated IDs, one ID as FILTER, INFO and FORMAT, lines in any order, generic and
structured lines, quoted commas, a missing fileformat line, samples; and on the lines the Broad operator appends."""
import glob
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, REPO

REF = os.path.join(REPO, "oracle", "_ref", "bin", "gdb_ref_dump")
CHROM = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref is not built")
    import __graft_entry__ as g
    lib = g.build_gpu_lib()
    exe = str(tmp_path_factory.mktemp("probe") / "header_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(REPO, "tests", "cpp", "header_probe.cc"), "-o", exe,
                           "-I" + os.path.join(REPO, "include"), "-L" + os.path.dirname(lib), "-lgdbgpu", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


TEMPLATE = open(os.path.join(GOLDEN, "inputs", "template_vcf_header.vcf")).read()
APPENDED = ["##INFO=<ID=NEW_INFO,Number=1,Type=Integer,Description=\"added\">",
            "##FORMAT=<ID=NEW_FMT,Number=R,Type=Float,Description=\"added, with a comma\">",
            "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"already there\">",
            "##FILTER=<ID=q10,Description=\"Quality below 10\">",
            "##contig=<ID=chrNew,length=12345>",
            "##contig=<ID=1,length=5>",
            "##source=a generic line",
            "SAMPLE HG00141", "SAMPLE HG01958", "SAMPLE HG00141", "SAMPLE with space"]
CASES = {
    "template": (TEMPLATE, None),
    "template_plus_appended_lines_and_samples": (TEMPLATE, APPENDED),
    "write_mode_start": ("", APPENDED),
    "no_fileformat_line": ("##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n" + CHROM, None),
    "idx_fields_out_of_order": ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\",IDX=0>\n##INFO=<ID=B,Number=1,Type=Integer,Description=\"b\",IDX=5>\n"
                                "##INFO=<ID=A,Number=1,Type=Integer,Description=\"a\",IDX=2>\n##FORMAT=<ID=C,Number=1,Type=Integer,Description=\"c\">\n"
                                "##contig=<ID=z,length=10,IDX=1>\n##contig=<ID=y,length=20,IDX=0>\n" + CHROM, ["##INFO=<ID=D,Number=1,Type=Integer,Description=\"d\">"]),
    "one_id_in_three_dictionaries": ("##fileformat=VCFv4.2\n##FILTER=<ID=DP,Description=\"f\">\n##INFO=<ID=DP,Number=1,Type=Integer,Description=\"i\">\n"
                                     "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"g\">\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"gt\">\n" + CHROM, None),
    "the_same_line_twice": ("##fileformat=VCFv4.2\n##INFO=<ID=DP,Number=1,Type=Integer,Description=\"first\">\n##INFO=<ID=DP,Number=1,Type=Integer,Description=\"second\">\n"
                            "##contig=<ID=1,length=10>\n##contig=<ID=1,length=20>\n" + CHROM, None),
    "no_pass_filter_and_other_filters_first": ("##fileformat=VCFv4.1\n##FILTER=<ID=LowQual,Description=\"Low quality\">\n##INFO=<ID=END,Number=1,Type=Integer,Description=\"e\">\n" + CHROM, None),
    "number_and_type_spellings": ("##fileformat=VCFv4.2\n##INFO=<ID=A,Number=A,Type=Float,Description=\"a\">\n##INFO=<ID=G,Number=G,Type=Integer,Description=\"g\">\n"
                                  "##INFO=<ID=R,Number=R,Type=Character,Description=\"r\">\n##INFO=<ID=V,Number=.,Type=String,Description=\"v\">\n"
                                  "##INFO=<ID=F,Number=0,Type=Flag,Description=\"f\">\n##FORMAT=<ID=N,Number=12,Type=Integer,Description=\"n\">\n" + CHROM, None),
    "extra_keys_quotes_and_commas": ("##fileformat=VCFv4.2\n##INFO=<ID=X,Number=1,Type=Integer,Description=\"has, commas = and \\\"quotes\\\"\",Source=\"s\",Version=\"3\">\n"
                                     "##ALT=<ID=NON_REF,Description=\"Represents any possible alternative allele at this location\">\n"
                                     "##GATKCommandLine=<ID=HaplotypeCaller,Version=3.4,Date=\"Mon\",CommandLineOptions=\"a=b c=d\">\n"
                                     "##contig=<ID=1,length=249250621,assembly=b37>\n##reference=file:///ref.fasta\n##source=SelectVariants\n" + CHROM, None),
    "samples_in_the_template": ("##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"gt\">\n" + CHROM[:-1] + "\tFORMAT\tS1\tS2\n", ["SAMPLE S3", "SAMPLE S1"]),
    "lines_after_the_contigs": ("##fileformat=VCFv4.2\n##contig=<ID=2,length=10>\n##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n##contig=<ID=1,length=10>\n"
                                "##FILTER=<ID=q,Description=\"q\">\n" + CHROM, ["##contig=<ID=3,length=3>", "##INFO=<ID=AF,Number=A,Type=Float,Description=\"af\">"]),
    # lines htslib drops, and headers it fails on
    "contig_lengths_that_are_no_numbers": ("##fileformat=VCFv4.2\n##contig=<ID=1>\n##contig=<ID=2,length=abc>\n##contig=<ID=3,length=12x>\n" + CHROM, None),
    "two_ids_with_one_idx": ("##fileformat=VCFv4.2\n##INFO=<ID=A,Number=1,Type=Integer,Description=\"a\",IDX=3>\n##INFO=<ID=B,Number=1,Type=Integer,Description=\"b\",IDX=3>\n" + CHROM, None),
    "two_contigs_with_one_idx": ("##fileformat=VCFv4.2\n##contig=<ID=a,length=1,IDX=2>\n##contig=<ID=b,length=1,IDX=2>\n" + CHROM, None),
    "a_sample_named_twice_in_the_template": ("##fileformat=VCFv4.2\n" + CHROM[:-1] + "\tFORMAT\tS1\tS2\tS1\n", None),
    "malformed_lines": ("##fileformat=VCFv4.2\n##INFO=<ID=X,Number=1,Type=Integer,Description=\"x\"\n##INFO=<Number=1,Type=Integer,Description=\"no id\">\nnot a header line\n"
                        "##INFO=<ID=Y, Number=1, Type=Double, Description=unquoted text>\n##FILTER=<ID=PASS,Description=\"mine\">\n##FILTER=<ID=q,Description=\"q\",IDX=9>\n" + CHROM, None),
    "no_sample_line_at_all": ("##fileformat=VCFv4.2\n##INFO=<ID=X,Number=1,Type=Integer,Description=\"x\">\n", None),
    "other_structured_lines": ("##fileformat=VCFv4.2\n##ALT=<ID=DEL,Description=\"d\">\n##PEDIGREE=<Name_0=a,Name_1=b>\n##SAMPLE=<ID=S,Genomes=g>\n##META=<ID=m,Type=String,Number=.,Values=[a, b]>\n" + CHROM, None),
    "windows_line_ends_are_not_special": ("##fileformat=VCFv4.2\n##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n\n" + CHROM, None),
}


@pytest.mark.parametrize("name", list(CASES))
def test_header_as_htslib_keeps_it(probe, tmp_path, name):
    text, lines = CASES[name]
    hf = tmp_path / "h.vcf"
    hf.write_text(text)
    ref_args, our_args = [REF, "--mode", "header", "-j", str(hf)], [probe, str(hf)]
    if lines is not None:
        lf = tmp_path / "lines.txt"
        lf.write_text("\n".join(lines) + "\n")
        ref_args += ["-l", str(lf)]
        our_args.append(str(lf))
    ref = subprocess.run(ref_args, capture_output=True, text=True)
    ours = subprocess.run(our_args, capture_output=True, text=True)
    assert (ref.returncode == 0) == (ours.returncode == 0), (ref.stderr[-400:], ours.stderr[-400:])
    if ref.returncode == 0:
        assert ours.stdout == ref.stdout


# ---- VCFAdapter::add_field_to_hdr_if_missing (vcf_adapter.cc:59-260): the line a queried field gets when the template lacks it
# (Number / Type from the vid mapping, "vcf_type" when given, multi-dimensional fields as Number=1 String keeping the template's
# Description, GT's two Descriptions), and the refusal when template and vid mapping disagree on a field's length or type
FIELD_VID = {"contigs": {"1": {"length": 1000, "tiledb_column_offset": 0}}, "fields": {
    "GT": {"type": "int", "length": "PP", "vcf_field_class": ["FORMAT"]},
    "PL": {"type": "int", "length": "G", "vcf_field_class": ["FORMAT"]},
    "AD": {"type": "int", "length": "R", "vcf_field_class": ["FORMAT"]},
    "AF": {"type": "float", "length": "A", "vcf_field_class": ["INFO"]},
    "SB": {"type": "int", "length": 4, "vcf_field_class": ["FORMAT"]},
    "PID": {"type": "char", "length": "VAR", "vcf_field_class": ["FORMAT"]},
    "DB": {"type": "flag", "vcf_field_class": ["INFO"]},
    "C1": {"type": "char", "length": 1, "vcf_field_class": ["INFO"]},
    "PLOIDY": {"type": "int", "length": "P", "vcf_field_class": ["FORMAT"]},
    "DP": {"type": "int", "vcf_field_class": ["INFO", "FORMAT"]},
    "MQ": {"type": "float", "vcf_field_class": ["INFO"]},
    "MQ0": {"type": "int", "vcf_field_class": ["INFO"]},
    "LowQual": {"type": "int", "vcf_field_class": ["FILTER"]},
    "AS": {"type": "float", "length": ["R", "var"], "vcf_field_class": ["INFO"], "vcf_delimiter": ["|", ","], "VCF_field_combine_operation": "element_wise_sum"},
    "HS": {"type": ["float", "int"], "length": ["R", "var"], "vcf_field_class": ["INFO"], "vcf_delimiter": ["|", ","], "VCF_field_combine_operation": "histogram_sum"},
    "VT": {"type": "int", "vcf_type": "String", "length": "VAR", "vcf_field_class": ["INFO"]},
    "VAR_IN_VID": {"type": "int", "length": "VAR", "vcf_field_class": ["INFO"]},
    "FLAG1": {"type": "flag", "length": 1, "vcf_field_class": ["INFO"]}}}
FIELD_TEMPLATE = ("##fileformat=VCFv4.2\n##INFO=<ID=MQ,Number=1,Type=Float,Description=\"template MQ\">\n##INFO=<ID=AS,Number=.,Type=Float,Description=\"template AS, wrong type\">\n"
                  "##FORMAT=<ID=SB,Number=2,Type=Integer,Description=\"wrong length\">\n##FORMAT=<ID=AD,Number=R,Type=Float,Description=\"wrong type\">\n"
                  "##INFO=<ID=DB,Number=0,Type=Flag,Description=\"flag ok\">\n##FORMAT=<ID=PID,Number=1,Type=String,Description=\"string number 1 ok\">\n"
                  "##INFO=<ID=MQ0,Number=1,Type=Float,Description=\"int in the vid mapping\">\n##INFO=<ID=VAR_IN_VID,Number=3,Type=Integer,Description=\"fixed here\">\n"
                  "##INFO=<ID=FLAG1,Number=0,Type=Flag,Description=\"length 1 in the vid mapping\">\n##INFO=<ID=AF,Number=.,Type=Float,Description=\"any variable length will do\">\n" + CHROM)
FIELD_LINES = ["FIELD F GT", "FIELD F PL", "FIELD F AD", "FIELD I AF", "FIELD F SB", "FIELD F PID", "FIELD I DB", "FIELD I C1", "FIELD F PLOIDY", "FIELD I DP", "FIELD F DP", "FIELD I MQ",
               "FIELD I MQ0", "FIELD T LowQual", "FIELD I AS", "FIELD I HS", "FIELD I VT", "FIELD I VAR_IN_VID", "FIELD I FLAG1", "FIELD T not_in_the_vid_mapping"]


def run_fields(probe, tmp_path, template, lines, vid_path):
    hf, lf = tmp_path / "h.vcf", tmp_path / "lines.txt"
    hf.write_text(template)
    lf.write_text("\n".join(lines) + "\n")
    ref = subprocess.run([REF, "--mode", "header", "-j", str(hf), "-l", str(lf), "-o", vid_path], capture_output=True, text=True)
    ours = subprocess.run([probe, str(hf), str(lf), vid_path], capture_output=True, text=True)
    assert ref.returncode == 0 and ours.returncode == 0, (ref.stderr[-400:], ours.stderr[-400:])
    assert ours.stdout == ref.stdout
    return ours.stdout


def test_field_lines_and_refusals(probe, tmp_path):
    vf = tmp_path / "vid.json"
    vf.write_text(json.dumps(FIELD_VID))
    out = run_fields(probe, tmp_path, FIELD_TEMPLATE, FIELD_LINES, str(vf))
    verdicts = dict(zip([l.split()[2] + "/" + l.split()[1] for l in FIELD_LINES], out.splitlines()))
    assert verdicts["GT/F"] == "field 1" and verdicts["MQ/I"] == "field 0" and verdicts["AS/I"] == "field 1"
    for name in ("AD/F", "SB/F", "MQ0/I", "VAR_IN_VID/I"):
        assert verdicts[name].endswith("refused"), name
    assert "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\",Description=\"GT\"" in out
    assert "##INFO=<ID=AS,Number=1,Type=String,Description=\"template AS, wrong type\"" in out


@pytest.mark.parametrize("vid", sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "inputs", "vid*.json"))))
def test_every_field_of_the_golden_vid_mappings_against_the_template(probe, tmp_path, vid):
    path = os.path.join(GOLDEN, "inputs", vid)
    fields = json.load(open(path))["fields"]
    lines = ["FIELD %s %s" % ({"INFO": "I", "FORMAT": "F", "FILTER": "T"}[c], n) for n, d in fields.items() for c in d.get("vcf_field_class", [])]
    out = run_fields(probe, tmp_path, TEMPLATE, lines, path)
    assert "refused" not in out
