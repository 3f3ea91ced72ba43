"""Every "path/File.java" + ":first-last" citation of the reference in this repository's headers, docs and sources must
name a file that exists in the reference tree and a line range inside it.  Needs /root/reference (the authoring
container); skipped where the tree is absent (GPU boxes)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
CITATION = re.compile(r"([A-Za-z0-9_\-./]*[A-Za-z0-9_\-]+\.(?:java|py|proto)):(\d+)(?:-(\d+))?")
OWN_PREFIXES = ("tests/", "lzy_b200/", "tools/", "oracle/", "integration/", "host/", "utils/", "api/v1/b200/", "core/lzy.py")


def sources():
    yield from (ROOT / n for n in ("DESIGN.md", "INTEGRATION.md", "README.md"))
    for pattern, base in (("*.h", "include"), ("*.py", "lzy_b200"), ("*.c*", "lzy_b200/csrc"), ("*.h", "lzy_b200/csrc"),
                          ("*.*", "oracle"), ("*.py", "tests"), ("*.*", "integration")):
        yield from (p for p in (ROOT / base).rglob(pattern) if p.is_file() and "_build" not in p.parts and p.suffix != ".so")


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the authoring container")
def test_reference_citations_point_into_existing_files():
    by_name = {}
    for p in REF.rglob("*"):
        if p.is_file() and p.suffix in (".java", ".py", ".proto"):
            by_name.setdefault(p.name, []).append(p)
    own_names = {p.name for p in ROOT.rglob("*") if p.is_file() and ".git" not in p.parts}
    checked, problems = 0, []
    for src in sources():
        try:
            text = src.read_text()
        except UnicodeDecodeError:
            continue
        for m in CITATION.finditer(text):
            path, first, last = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            name = path.rsplit("/", 1)[-1]
            candidates = by_name.get(name, [])
            if not candidates:
                if name in own_names or path.startswith(OWN_PREFIXES):
                    continue  # a citation of this repository's own file
                problems.append(f"{src.relative_to(ROOT)}: {m.group(0)} — no such file in the reference")
                continue
            parts = [x for x in path.split("/") if x not in ("", ".", "..")]
            def score(c):
                return sum(1 for x in parts if x in c.parts or x == c.name)

            top = max(score(c) for c in candidates)
            # namesakes that match the cited path equally well (two graph executors ship an algo/Algorithms.java):
            # the range has to fit one of them
            matching = [c for c in candidates if score(c) == top]
            longest = max(sum(1 for _ in c.open(errors="ignore")) for c in matching)
            checked += 1
            if first > last or last > longest:
                problems.append(f"{src.relative_to(ROOT)}: {m.group(0)} — {matching[0].relative_to(REF)} has {longest} lines")
    assert checked > 250, checked
    assert not problems, "\n".join(problems)
