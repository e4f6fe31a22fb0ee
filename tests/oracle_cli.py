"""Test helper: parse a reference-style argv (bedops / bedmap / closest-features) into a neutral call description,
then run it either through the Python oracle (checker) or through the CUDA library's C ABI (product)."""
import bed_oracle as O

BEDMAP_OPS = {"echo", "count", "indicator", "bases", "sum", "mean", "max", "min", "echo-map-id", "echo-ref-size",
              "echo-ref-name", "echo-ref-row-id", "echo-map", "echo-map-score", "echo-map-size", "echo-overlap-size",
              "echo-map-range", "bases-uniq", "bases-uniq-f", "variance", "stdev", "cv", "echo-map-id-uniq", "median", "wmean",
              "max-element", "min-element"}


def parse_argv(tool, argv, known_files):
    argv = list(argv)
    d = dict(tool=tool, names=[], chrom=None)
    i = 0
    if tool == "bedops":
        d.update(op=None, thr=1.0, pct=True, full_left=False, pad=(0, 0))
        while i < len(argv):
            a = argv[i]
            if a == "--ec":
                pass
            elif a == "--chrom":
                i += 1
                d["chrom"] = argv[i].encode()
            elif a == "--range":                                  # Input.hpp:86-126
                i += 1
                v = argv[i]
                d["pad"] = tuple(int(x) for x in v.split(":")) if ":" in v else (-int(v), int(v))
            elif a in ("-p", "--partition"):
                d["op"] = "partition"
            elif a in ("-m", "--merge"):
                d["op"] = "merge"
            elif a in ("-i", "--intersect"):
                d["op"] = "intersect"
            elif a in ("-c", "--complement"):
                d["op"] = "complement"
            elif a == "-L":
                d["full_left"] = True
            elif a in ("-d", "--difference"):
                d["op"] = "difference"
            elif a in ("-s", "--symmdiff"):
                d["op"] = "symmdiff"
            elif a in ("-u", "--everything"):
                d["op"] = "everything"
            elif a in ("-w", "--chop"):
                d["op"] = "chop"
                d.update(chunk=1, stagger=0, exclude_short=False)
                while i + 1 < len(argv):                      # Input.hpp:221-258
                    n = argv[i + 1]
                    if n == "--stagger":
                        d["stagger"] = int(argv[i + 2])
                        i += 2
                    elif n == "-x":
                        d["exclude_short"] = True
                        i += 1
                    elif n.isdigit() and n not in known_files:
                        d["chunk"] = int(n)
                        i += 1
                    else:
                        break
            elif a in ("-e", "--element-of", "-n", "--not-element-of"):
                d["op"] = "element-of" if a in ("-e", "--element-of") else "not-element-of"
                if i + 1 < len(argv) and argv[i + 1] not in known_files:
                    s = argv[i + 1]
                    if s.endswith("%"):
                        d["thr"], d["pct"] = float(s[:-1].lstrip("-")) / 100.0, True
                        if d["thr"] == 0:
                            d["thr"], d["pct"] = 1.0, False
                    else:
                        d["thr"], d["pct"] = float(int(s.lstrip("-"))), False
                    i += 1
            else:
                d["names"].append(a)
            i += 1
    elif tool == "bedmap":
        d.update(ops=[], overlap=("bp", 1), prec=6, sci=False, delim=b"|", multidelim=b";", skip_unmapped=False)
        while i < len(argv):
            a = argv[i]
            if a.startswith("--") and a[2:] in BEDMAP_OPS:
                d["ops"].append(a[2:])
            elif a == "--kth":
                i += 1
                k = float(argv[i])                                # 0 -> min, 1 -> max (Bedmap.cpp:495-498)
                d["ops"].append("min" if k == 0 else "max" if k == 1 else "kth:" + argv[i])
            elif a == "--tmean":
                d["ops"].append("tmean:%s:%s" % (argv[i + 1], argv[i + 2]))
                i += 2
            elif a == "--mad":   # optional multiplier (Input.hpp:275-288): a following all-numeric argument
                if i + 1 < len(argv) and argv[i + 1] and all(c in ".-0123456789" for c in argv[i + 1]):
                    i += 1
                    d["ops"].append("mad:" + argv[i])
                else:
                    d["ops"].append("mad")
            elif a == "--prec":
                i += 1
                d["prec"] = int(argv[i])
            elif a == "--delim":
                i += 1
                d["delim"] = argv[i].encode()
            elif a == "--multidelim":
                i += 1
                d["multidelim"] = argv[i].encode()
            elif a == "--chrom":
                i += 1
                d["chrom"] = argv[i].encode()
            elif a == "--skip-unmapped":
                d["skip_unmapped"] = True
            elif a == "--sci":
                d["sci"] = True
            elif a in ("--bp-ovr", "--range"):
                i += 1
                d["overlap"] = ("bp" if a == "--bp-ovr" else "range", int(argv[i]))
                if d["overlap"] == ("range", 0):
                    d["overlap"] = ("bp", 1)
            elif a.startswith("--fraction-"):
                i += 1
                d["overlap"] = (a[2:], float(argv[i]))
            elif a == "--exact":
                d["overlap"] = ("exact", 0)
            elif a in ("--ec", "--faster", "--sweep-all"):
                pass
            else:
                d["names"].append(a)
            i += 1
    elif tool == "closest-features":
        d.update(dist=False, closest=False, no_overlaps=False, no_ref=False, delim=b"|")
        while i < len(argv):
            a = argv[i]
            if a == "--dist":
                d["dist"] = True
            elif a in ("--closest", "--shortest"):
                d["closest"] = True
            elif a == "--no-overlaps":
                d["no_overlaps"] = True
            elif a == "--no-ref":
                d["no_ref"] = True
            elif a == "--delim":
                i += 1
                d["delim"] = argv[i].encode()
            elif a == "--chrom":
                i += 1
                d["chrom"] = argv[i].encode()
            else:
                d["names"].append(a)
            i += 1
    else:
        raise ValueError(tool)
    return d


def run(tool, argv, files, stdin=None):
    """Oracle leg.  files: {name: bytes}.  Returns stdout bytes."""
    d = parse_argv(tool, argv, files)
    texts = [stdin if n == "-" else files[n] for n in d["names"]]
    if tool == "bedops":
        if d["pad"] != (0, 0):   # the operators read the padded view; -e/-n leave the reference file alone
            first = 1 if d["op"].endswith("element-of") else 0
            padded = [O.pad_rows(O.parse_bed(t, 3), *d["pad"]) for t in texts[first:]]
            if d["op"] == "everything":
                return O.merge_everything([O._sel(f, d["chrom"]) for f in padded])
            texts = texts[:first] + [O._rows_to_text(f) for f in padded]
        if d["op"] == "partition":
            return O.bedops_partition(texts, d["chrom"])
        if d["op"] == "merge":
            return O.bedops_merge(texts, d["chrom"])
        if d["op"] == "intersect":
            return O.bedops_intersect(texts, d["chrom"])
        if d["op"] == "complement":
            return O.bedops_complement(texts, d["full_left"], d["chrom"])
        if d["op"] == "difference":
            return O.bedops_difference(texts, d["chrom"])
        if d["op"] == "symmdiff":
            return O.bedops_symmdiff(texts, d["chrom"])
        if d["op"] == "everything":
            return O.bedops_everything(texts, d["chrom"])
        if d["op"] == "chop":
            return O.bedops_chop(texts, d["chunk"], d["stagger"], d["exclude_short"], d["chrom"])
        return O.bedops_element_of(texts, d["thr"], d["pct"], d["op"] == "not-element-of", d["chrom"])
    if tool == "bedmap":
        return O.bedmap(texts[0], texts[1] if len(texts) > 1 else None, ops=d["ops"], overlap=d["overlap"],
                        prec=d["prec"], sci=d["sci"], delim=d["delim"], multidelim=d["multidelim"],
                        skip_unmapped=d["skip_unmapped"], chrom=d["chrom"])
    return O.closest_features(texts[0], texts[1], dist=d["dist"], closest=d["closest"], no_overlaps=d["no_overlaps"],
                              no_ref=d["no_ref"], delim=d["delim"], chrom=d["chrom"])


def run_kit(kit, tool, argv, files, stdin=None):
    """Product leg: the same call through libbedkit.so's C ABI (ctypes).  Mirrors what the C++ tools do."""
    from bedops_b200._lib import COL_ID, COL_LINE, COL_SCORE
    d = parse_argv(tool, argv, files)
    texts = [stdin if n == "-" else files[n] for n in d["names"]]
    if tool == "bedops":
        beds = [kit.load(t, 3, COL_LINE if ((k == 0 and d["op"].endswith("element-of")) or d["op"] == "everything") else 0)
                for k, t in enumerate(texts)]
        thr = float(d["full_left"]) if d["op"] == "complement" else d["thr"]   # BK_SETOP_COMPLEMENT: thr != 0 is -L
        if d["pad"] != (0, 0):
            first = 1 if d["op"].endswith("element-of") else 0
            beds = beds[:first] + [kit.pad(b, *d["pad"]) for b in beds[first:]]
        if d["op"] == "chop":
            out = kit.chop(beds, d["chunk"], d["stagger"], d["exclude_short"], d["chrom"])
        else:
            out = kit.setop(d["op"], beds, thr, d["pct"], d["chrom"])
        for b in beds:
            b.free()
        return out
    if tool == "bedmap":
        ops = d["ops"]
        score = any(o in ("sum", "mean", "max", "min", "echo-map-score", "variance", "stdev", "cv", "median", "wmean", "max-element",
                          "min-element") or o.startswith("kth:") or o.startswith("mad") or o.startswith("tmean:") for o in ops)
        element = any(o in ("max-element", "min-element") for o in ops)
        ids = "echo-map-id" in ops or "echo-map-id-uniq" in ops
        line = any(o in ("echo", "echo-ref-name", "echo-map-range") for o in ops)
        mf = 5 if score else (4 if ids else 3)
        mcols = (COL_SCORE if score else 0) | ((COL_ID | COL_LINE) if ids else 0) | (COL_LINE if "echo-map" in ops or element else 0)
        if len(texts) > 1:
            ref = kit.load(texts[0], 3, COL_LINE if line else 0)
            mp = kit.load(texts[1], mf, mcols)
        else:
            ref = kit.load(texts[0], mf, mcols | (COL_LINE if line else 0))
            mp = None
        out = kit.bedmap(ref, mp, ops, overlap=d["overlap"], prec=d["prec"], sci=d["sci"], delim=d["delim"],
                         multidelim=d["multidelim"], skip_unmapped=d["skip_unmapped"], chrom=d["chrom"])
        ref.free()
        if mp is not None:
            mp.free()
        return out
    ref = kit.load(texts[0], 3, COL_LINE)
    qry = kit.load(texts[1], 3, COL_LINE)
    out = kit.closest(ref, qry, dist=d["dist"], closest=d["closest"], no_overlaps=d["no_overlaps"], no_ref=d["no_ref"],
                      delim=d["delim"], chrom=d["chrom"])
    ref.free()
    qry.free()
    return out
