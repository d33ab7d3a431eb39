#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference checkout.

Run in the build container (the GPU box has no /root/reference; the committed JSON files are what
travels):    python tests/golden/make_golden.py [/root/reference]

Fixtures written
  th_indexing_variants.json   the reference manual's table of the multi-indices of the eight
                              Taylor-Hood index-merging variants, parsed literally from
                              doc/manual/dune-functions-bases.tex (Table tab:th_indexing_variants);
                              entries stay symbolic in n_2 = dim(Q2), e.g. "(0,2n_2+1)".
  merging_strategies.json     the four doc-comment tables of
                              dune/functions/functionspacebases/basistags.hh (what each index
                              merging strategy does to two children {f}, {g}), parsed literally.
  face_orientations.json      the truth tables of Experimental::FaceOrientations in the doc comment of
                              dune/functions/functionspacebases/lagrangebasis.hh (:132-174): for every
                              ordering of the global vertex ids of a quadrilateral (and triangle) the
                              edge orientations and the orientation index, parsed literally ("^" rows
                              inherit the index of the row above, as in the comment).
  numbering_examples.json     the hand-derived numbering vectors of SURVEY.md Appendix A (A1-A3);
                              these are NOT from the reference (they document the dune-grid
                              conventions the oracle assumes) and are marked as such.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def parse_th_table():
    path = os.path.join(REF, "doc/manual/dune-functions-bases.tex")
    lines = open(path, encoding="utf-8").read().split("\n")
    end = next(i for i, l in enumerate(lines) if "label{tab:th_indexing_variants}" in l)
    begin = max(i for i, l in enumerate(lines[:end]) if "\\begin{tabular}" in l)
    body = "\n".join(l for l in lines[begin + 1:end] if not l.lstrip().startswith("%"))
    rows = [r.strip() for r in body.split("\\\\")]
    header = [c.strip() for c in rows[0].split("&")][1:]
    assert header == ["BL(BL)", "BL(BI)", "BL(FL)", "BL(FI)", "FL(BL)", "FL(BI)", "FL(FL)", "FL(FI)"], header
    table = {h: {} for h in header}
    for r in rows[1:]:
        r = r.replace("\\hline", "").strip()
        cells = [c.strip().strip("$") for c in r.split("&")]
        if len(cells) != 9 or "vdots" in cells[0]:
            continue
        m = re.fullmatch(r"v_\{x_(\d),(\d)\}", cells[0])
        key = f"v{m.group(1)},{m.group(2)}" if m else "p" + re.fullmatch(r"p_\{(\d)\}", cells[0]).group(1)
        for h, c in zip(header, cells[1:]):
            table[h][key] = c
    return {"source": f"doc/manual/dune-functions-bases.tex:{begin + 1}-{end + 1}",
            "note": "row v<c>,<i> = velocity component c, i-th Q2 basis function; p<j> = j-th Q1 function; "
                    "n_2 = dimension of the Q2 basis; the table is for 3 velocity components",
            "columns": table}


def parse_merging_tables():
    path = os.path.join(REF, "dune/functions/functionspacebases/basistags.hh")
    text = open(path, encoding="utf-8").read()
    out = {}
    for name in ("FlatLexicographic", "FlatInterleaved", "BlockedLexicographic", "BlockedInterleaved"):
        stop = text.index(f"struct {name}")
        start = text.rindex("/**", 0, stop)
        doc = text[start:stop]
        tabs = {}
        cur = None
        for line in doc.split("\n"):
            line = line.strip().lstrip("*").strip()
            m = re.match(r"function in \{([fg,]+)\}\|\s*index", line)
            if m:
                cur = m.group(1)
                tabs[cur] = []
                continue
            m = re.match(r"([fg])_(\d)\s*\|\s*\(([^)]*)\)", line)
            if m and cur:
                tabs[cur].append([m.group(1) + m.group(2), [x.strip() for x in m.group(3).split(",")]])
        assert set(tabs) == {"f", "g", "f,g"} and len(tabs["f,g"]) == 6, (name, tabs)
        out[name] = {"f": tabs["f"], "g": tabs["g"], "merged": tabs["f,g"]}
    return {"source": "dune/functions/functionspacebases/basistags.hh:52-184", "tables": out}


def parse_face_orientation_tables():
    path = os.path.join(REF, "dune/functions/functionspacebases/lagrangebasis.hh")
    lines = open(path, encoding="utf-8").read().split("\n")
    out = {"source": "dune/functions/functionspacebases/lagrangebasis.hh doc comment of Experimental::FaceOrientations",
           "triangle": [], "quadrilateral": []}
    section = None
    last = None
    for ln in lines:
        if "**Triangle orientations**" in ln:
            section = "triangle"
        elif "**Quadrilateral orientations**" in ln:
            section = "quadrilateral"
        elif "**Complexity bounds**" in ln:
            break
        if section is None or "|" not in ln:
            continue
        cells = [c.strip() for c in ln.strip().lstrip("*").strip().strip("|").split("|")]
        if len(cells) != 9 or not re.fullmatch(r"[\d,]+", cells[0]):
            continue
        ids = [int(x) for x in cells[0].split(",")]
        edges = [int(x) for x in cells[1].split(",")]
        if cells[8] == "^":
            row = dict(last, vertex_ids=ids, edge_orientations=edges)
        else:
            row = {"vertex_ids": ids, "edge_orientations": edges, "primary_edge": cells[2],
                   "rotations": int(cells[3]), "reflection": cells[4] == "yes", "bits": cells[7], "index": int(cells[8])}
        last = {k: v for k, v in row.items() if k not in ("vertex_ids", "edge_orientations")}
        out[section].append(row)
    assert len(out["triangle"]) == 6 and len(out["quadrilateral"]) == 24, (len(out["triangle"]), len(out["quadrilateral"]))
    return out


def numbering_examples():
    return {
        "source": "SURVEY.md Appendix A (hand-derived from lagrangebasis.hh:843-873 plus the recalled dune-grid "
                  "conventions; NOT reference output - they pin the oracle's documented assumptions)",
        "A1_q2_2d_2x1": {"grid": [2, 1], "order": 2, "dimension": 15,
                         "element_indices": [[9, 2, 10, 6, 0, 7, 12, 4, 13], [10, 3, 11, 7, 1, 8, 13, 5, 14]]},
        "A2_q2_2d_1x1": {"grid": [1, 1], "order": 2, "dimension": 9, "element_indices": [[5, 1, 6, 3, 0, 4, 7, 2, 8]]},
        "A2_th_flfl_2d_1x1": {"grid": [1, 1], "dimension": 22,
                              "element_indices": [[5, 1, 6, 3, 0, 4, 7, 2, 8, 14, 10, 15, 12, 9, 13, 16, 11, 17,
                                                   18, 19, 20, 21]]},
        "A3_q2_3d_1x1x1": {"grid": [1, 1, 1], "order": 2, "dimension": 27,
                           "element_indices": [[19, 7, 20, 11, 1, 12, 21, 8, 22, 15, 3, 16, 5, 0, 6, 17, 4, 18,
                                                23, 9, 24, 13, 2, 14, 25, 10, 26]]},
    }


def main():
    global REF
    if len(sys.argv) > 1:
        REF = sys.argv[1]
    for name, data in (("th_indexing_variants.json", parse_th_table()),
                       ("merging_strategies.json", parse_merging_tables()),
                       ("face_orientations.json", parse_face_orientation_tables()),
                       ("numbering_examples.json", numbering_examples())):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
            f.write("\n")
        print("wrote", name)


if __name__ == "__main__":
    main()
