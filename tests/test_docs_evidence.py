"""Documentation hygiene (CPU): every evidence file that profiles/README.md, DESIGN.md or INTEGRATION.md names under
profiles/ exists in the tree, so a cited number can always be traced to its capture."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _named_files(text):
    names = set()
    for m in re.finditer(r"`(?:profiles/)?((?:r[12]_|ncu_|sass_)[A-Za-z0-9_.{},|\-]+\.(?:json|jsonl|csv|log|txt))`", text):
        name = m.group(1)
        brace = re.search(r"\{([^}]*)\}", name)
        if brace:  # `r2_search_kernel_{f32,int8}_raw.csv` style lists
            for alt in re.split(r"[,|]", brace.group(1)):
                names.add(name[:brace.start()] + alt + name[brace.end():])
        else:
            names.add(name)
    return {n for n in names if "{" not in n and "…" not in n}


def test_every_cited_profile_file_exists():
    missing = {}
    for doc in ("profiles/README.md", "DESIGN.md", "INTEGRATION.md", "README.md"):
        text = open(os.path.join(ROOT, doc)).read()
        gone = sorted(n for n in _named_files(text) if not os.path.exists(os.path.join(ROOT, "profiles", n)))
        if gone:
            missing[doc] = gone
    assert not missing, missing
