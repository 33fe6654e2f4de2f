"""CPU: the committed fixtures under tests/golden/ are exactly what their generator writes (tests/golden/make_fixtures.py: hand
transcriptions of the inputs and assertions of the reference's own tests, each citing file and lines), and every line range a fixture
cites exists in the reference checkout when there is one (the build container)."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def test_fixtures_are_what_the_generator_writes(tmp_path):
    shutil.copy(os.path.join(GOLDEN, "make_fixtures.py"), tmp_path / "make_fixtures.py")
    r = subprocess.run([sys.executable, str(tmp_path / "make_fixtures.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    made = sorted(f for f in os.listdir(tmp_path) if f.endswith(".json"))
    assert made == sorted(f for f in os.listdir(GOLDEN) if f.endswith(".json")), made
    for f in made:
        assert json.load(open(tmp_path / f)) == json.load(open(os.path.join(GOLDEN, f))), f"{f}: committed fixture differs from the generator's output"


def test_cited_reference_lines_exist():
    if not os.path.isdir("/root/reference"):
        pytest.skip("no reference checkout on this machine")
    cited = 0
    for f in sorted(os.listdir(GOLDEN)):
        if not f.endswith(".json"):
            continue
        src = json.load(open(os.path.join(GOLDEN, f))).get("source", "")
        for path, spans in re.findall(r"(/root/reference/[\w/\.\-]+\.rs):([\d\-,: ]+)", src):
            assert os.path.isfile(path), (f, path)
            n_lines = sum(1 for _ in open(path, errors="replace"))
            for a in re.findall(r"\d+", spans):
                assert 1 <= int(a) <= n_lines, (f, path, a, n_lines)
                cited += 1
    assert cited >= 10


def test_source_citations_resolve():
    """every `file.rs:line[-line]` / `file.cu:line` citation in the repository's sources and documents names a file of the reference
    checkout (by path suffix) that has that many lines — a mistyped or stale citation fails here"""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("no reference checkout on this machine")
    root = os.path.dirname(HERE)
    ref_files = {}
    for d, _, fs in os.walk(ref):
        for f in fs:
            if f.endswith((".rs", ".cu")):
                p = os.path.join(d, f)
                ref_files[p] = sum(1 for _ in open(p, errors="replace"))
    listed = subprocess.run(["git", "ls-files"], cwd=root, stdout=subprocess.PIPE, text=True).stdout.split()
    ours = {"lib.rs", "ffi.rs", "planner.rs", "reasoner.rs", "r2r.rs"}  # rust_shim's own files
    skip = ("SURVEY", "BASELINE", "PAPERS", "SNIPPETS", "VERDICT", "ADVICE")
    total, bad = 0, []
    for s in listed:
        if not s.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".md", ".cpp", ".rs", ".sh")) or s.startswith(skip):
            continue
        text = open(os.path.join(root, s), errors="replace").read()
        for m in re.finditer(r"([A-Za-z_][\w/\.\-]*\.(?:rs|cu)):(\d+)(?:-(\d+))?", text):
            path, last = m.group(1), int(m.group(3) or m.group(2))
            path = path[len(ref) + 1:] if path.startswith(ref + "/") else path
            if path.startswith(("kb_", "rust_shim", "tests/", "kolibrie_b200/")) or (path in ours and "rust_shim" in s):
                continue
            cands = [n for f, n in ref_files.items() if f.endswith("/" + path)]
            total += 1
            if not cands or max(cands) < last:
                bad.append((s, m.group(0)))
    assert total >= 300 and not bad, bad[:20]
