"""TEST INFRASTRUCTURE ONLY (oracle): a pure-Python restatement of `cramore dsc-pileup`
(cmd_cram_dsc_pileup.cpp) and of the BAM x VCF join it shares with `cramore demuxlet --sam`
(cmd_cram_demuxlet.cpp:222-395), written from the reference's control flow with Python dicts standing in for its
std::maps.  Only tests/ may import it; the product (cramore_b200/host/bam_io.h, cmd_dsc_pileup.cpp) never does.

PARITY UNPINNED: the reference's own command needs htslib and cannot be built in this environment, so this file is a
second, independent reading of the same source lines, not a recording of the reference's output.

Followed lines:
  * read filter                      sam_filtered_reader.cpp:298-310 (--min-MQ, --excl-flag), dsc-pileup :208-212
  * VCF cursor / SNP window          cmd_cram_dsc_pileup.cpp:222-246, bcf_filtered_reader.cpp:722-742 (clear_buffer_before)
  * droplet / UMI tags               :249-293
  * covered regions per UMI          :296-331, genomeLoci.h:17-75 (ordering), :227-237 (add), :255-287 (resolveOverlaps)
  * base at a SNP                    hts_utils.cpp:279-358
  * first read of (droplet,SNP,UMI)  sc_drop_seq.cpp:45-91
  * output files                     cmd_cram_dsc_pileup.cpp:438-523
"""
import gzip
import struct

NT16 = "=ACMGRSVTWYHKDBN"
CIGAR_OPS = "MIDNSHP=XB"


def read_bam(path):
    """-> (contig names, records).  BGZF is a series of gzip members, which gzip.open reads as one stream."""
    with gzip.open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", data, 4)
    off = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, off)
    off += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, off)
        refs.append(data[off + 4:off + 4 + l_name].rstrip(b"\0").decode())
        off += 4 + l_name + 4
    recs = []
    while off < len(data):
        block_size, = struct.unpack_from("<i", data, off)
        b = data[off + 4:off + 4 + block_size]
        off += 4 + block_size
        tid, pos, l_name, mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", b, 0)
        p = 32 + l_name
        cigar = []
        for i in range(n_cigar):
            v, = struct.unpack_from("<I", b, p + 4 * i)
            cigar.append((CIGAR_OPS[v & 15], v >> 4))
        p += 4 * n_cigar
        seq = "".join(NT16[(b[p + (i >> 1)] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        p += (l_seq + 1) // 2
        qual = list(b[p:p + l_seq])
        p += l_seq
        tags = {}
        while p + 3 <= len(b):
            tag, typ = b[p:p + 2].decode(), chr(b[p + 2])
            p += 3
            if typ in "Z":
                e = b.index(b"\0", p)
                tags[tag] = b[p:e].decode()
                p = e + 1
            elif typ in "AcC":
                p += 1
            elif typ in "sS":
                p += 2
            elif typ in "iIf":
                p += 4
            else:
                raise ValueError("aux type %r not handled by this test reader" % typ)
        recs.append(dict(tid=tid, pos=pos, mapq=mapq, flag=flag, cigar=cigar, seq=seq, qual=qual, tags=tags, l_seq=l_seq))
    return refs, recs


def read_vcf(path):
    """-> (contig names of the header, records [dict(chrom,pos,ref,alt,info)])"""
    opener = gzip.open if path.endswith(".gz") else open
    contigs, out = [], []
    with opener(path, "rt") as f:
        for ln in f:
            if ln.startswith("##contig=<ID="):
                contigs.append(ln[13:].split(",")[0].split(">")[0].strip())
            if ln.startswith("#") or not ln.strip():
                continue
            t = ln.rstrip("\n").split("\t")
            info = dict(kv.split("=", 1) for kv in t[7].split(";") if "=" in kv)
            out.append(dict(chrom=t[0], pos=int(t[1]), ref=t[3], alt=t[4], info=info, cols=t[9:], fmt=t[8] if len(t) > 8 else ""))
    return contigs, out


def endpos(r):
    return r["pos"] + sum(n for op, n in r["cigar"] if op in "MDN=X")


def base_at(r, pos):
    """hts_utils.cpp:279-358 -> (base, Phred+33 quality, read index or None)"""
    rlen, cpos, rpos = r["l_seq"], r["pos"], 0
    base, qual = "N", 0
    if r["cigar"]:
        for op, ln in r["cigar"]:
            if op == "M":
                if cpos <= pos <= cpos + ln - 1:
                    rpos += pos - cpos
                    break
                cpos += ln
                rpos += ln
            elif op in "DN":
                if cpos <= pos <= cpos + ln - 1:
                    rpos = -1
                    break
                cpos += ln
            elif op in "SI":
                rpos += ln
        if 0 <= rpos <= rlen:
            base = r["seq"][rpos] if rpos < rlen else "\0"
            qual = r["qual"][rpos] + 33 if rpos < rlen else 0
        else:
            rpos = None
    if rpos is not None and rpos >= rlen:
        rpos, base = None, "."
    return base, qual, rpos


def _locus_less(names):
    def less(a, b):  # genomeLocus::operator<; a, b = (contig name, beg1, end0)
        if a[0] == b[0]:
            return a[2] < b[2] if a[1] == b[1] else a[1] < b[1]

        def atoi(s):
            n = 0
            for ch in s.lstrip():
                if not ch.isdigit():
                    break
                n = n * 10 + int(ch)
            return n
        n1, n2 = atoi(a[0]), atoi(b[0])
        if n1 == 0 and n2 == 0:
            return a[0] < b[0]
        if n1 > 0 and n2 > 0:
            return n1 < n2
        return n1 > 0
    return less


class Loci:
    """genomeLoci on a Python list kept in the order of the reference's std::set"""

    def __init__(self, less):
        self.v, self.less = [], less

    def _find(self, x):
        for i, y in enumerate(self.v):
            if not self.less(y, x):
                return i
        return len(self.v)

    def add(self, x):
        i = self._find(x)
        if i < len(self.v) and not self.less(x, self.v[i]):
            return i  # an equivalent locus is in the set
        self.v.insert(i, x)
        return i

    def resolve(self):
        i = 1
        while i < len(self.v):
            p, c = self.v[i - 1], self.v[i]
            if p[0] == c[0] and p[1] <= c[2] and c[1] <= p[2]:
                m = (p[0], min(p[1], c[1]), max(p[2], c[2]))
                del self.v[i - 1:i + 1]
                i = self.add(m) + 1
            else:
                i += 1


def dsc_pileup(bam_path, vcf_path, tag_group="CB", tag_umi="UB", cap_bq=40, min_bq=13, min_mq=20, min_td=0,
               excl_flag=0x0f04, exclude_flag=0x0704, group_list=None, skip_umi=False, demuxlet_mode=False):
    """-> dict(cel=text, var=text, plp=text, umi=text or None, store=...) with the exact file contents."""
    refs, recs = read_bam(bam_path)
    contigs, variants = read_vcf(vcf_path)
    for v in variants:  # htslib adds contigs missing from the header as it meets them
        if v["chrom"] not in contigs:
            contigs.append(v["chrom"])
    rid_of = {c: i for i, c in enumerate(contigs)}
    rchroms = {}
    prevrid = -1
    for t, name in enumerate(refs):
        rid = rid_of.get(name, -1)
        if rid >= 0:
            assert prevrid < rid, "contig order differs"
            rchroms[rid] = name
            prevrid = rid
    assert rchroms

    def af_of(v):
        if "AF" in v["info"]:
            return float(struct.unpack("f", struct.pack("f", float(v["info"]["AF"].split(",")[0])))[0])
        return int(v["info"]["AC"].split(",")[0]) / int(v["info"]["AN"])

    snps = []  # dict(rid,pos,ref,alt,af,rlen)
    nxt = [0]

    def vcf_read():
        if nxt[0] >= len(variants):
            return False
        v = variants[nxt[0]]
        nxt[0] += 1
        snps.append(dict(rid=rid_of[v["chrom"]], pos=v["pos"], ref=v["ref"][0], alt=v["alt"][0], rlen=len(v["ref"]),
                         af=None if demuxlet_mode else af_of(v)))
        return True

    assert vcf_read()
    eof = False
    ibeg = 0
    cells, totl, uniq = {}, [], []
    umi_loci = []   # per droplet: {umi: (Loci fwd, Loci rev)}
    snp_umis = {}   # snp -> {cell: {umi: (al, bq)}}
    less = _locus_less(refs)
    for r in recs:
        if r["mapq"] < min_mq or (r["flag"] & excl_flag):
            continue
        ep = endpos(r)
        chrom = refs[r["tid"]] if 0 <= r["tid"] < len(refs) else None
        tid2rid = rid_of.get(chrom, -1)
        no_bcf = False
        if demuxlet_mode:
            if tid2rid < 0:
                continue
        else:
            if r["flag"] & exclude_flag:
                continue
            no_bcf = tid2rid < 0
        crid = snps[-1]["rid"]
        while ibeg < len(snps):
            v = snps[ibeg]
            if v["rid"] < crid or (v["rid"] == crid and (v["pos"] - 1) + v["rlen"] < r["pos"]):
                ibeg += 1
            else:
                break
        if not no_bcf:
            while not eof and (snps[-1]["rid"] < tid2rid or (snps[-1]["rid"] == tid2rid and snps[-1]["pos"] - 1 < ep)):
                if not vcf_read():
                    eof = True
        sbcd = "."
        if tag_group:
            sbcd = r["tags"].get(tag_group, ".")
            if group_list is not None and sbcd not in group_list:
                continue
        if sbcd not in cells:
            cells[sbcd] = len(cells)
            totl.append(0)
            uniq.append(0)
            umi_loci.append({})
        ibcd = cells[sbcd]
        sumi = r["tags"].get(tag_umi, ".")
        totl[ibcd] += 1
        if not skip_umi and not demuxlet_mode:
            pair = umi_loci[ibcd].setdefault(sumi, (Loci(less), Loci(less)))
            loci = pair[1] if r["flag"] & 0x10 else pair[0]
            if r["cigar"]:
                cpos = r["pos"]
                for op, ln in r["cigar"]:
                    if op == "M":
                        loci.add((chrom, cpos + 1, cpos + ln))
                        cpos += ln
                    elif op in "DN":
                        cpos += ln
                loci.resolve()
        if no_bcf:
            continue
        for i in range(ibeg, len(snps)):
            base, qual, rpos = base_at(r, snps[i]["pos"] - 1)
            if rpos is None or base == "N":
                continue
            if qual - 33 < min_bq or rpos < min_td - 1 or rpos + min_td > r["l_seq"]:
                continue
            allele = 0 if base == snps[i]["ref"] else (1 if base == snps[i]["alt"] else 2)
            bq = min(qual - 33, cap_bq)
            d = snp_umis.setdefault(i, {}).setdefault(ibcd, {})
            if sumi not in d:
                d[sumi] = (allele, bq)
                uniq[ibcd] += 1
    # ---- files ----
    bcs = sorted(cells, key=cells.get)
    nsnp_cell = [0] * len(bcs)
    for i, per_cell in snp_umis.items():
        for c in per_cell:
            nsnp_cell[c] += 1
    cel = ["#DROPLET_ID\tBARCODE\tNUM.READ\tNUM.UMI\tNUM.UMIwSNP\tNUM.SNP\n"]
    for i, b in enumerate(bcs):
        numi = "." if skip_umi else "%d" % len(umi_loci[i])
        cel.append("%d\t%s\t%d\t%s\t%d\t%d\n" % (i, b, totl[i], numi, uniq[i], nsnp_cell[i]))
    umi = None
    if not skip_umi:
        umi = []
        for i in range(len(bcs)):
            for u in sorted(umi_loci[i]):
                fw, rv = umi_loci[i][u]
                if not fw.v and not rv.v:
                    continue
                row = "%d\t%s\t%d\t%d" % (i, u, sum(e - b + 1 for _, b, e in fw.v), sum(e - b + 1 for _, b, e in rv.v))
                for (c, b, e) in fw.v:
                    row += "\t%s:%d:%d:+" % (c, b, e - b + 1)
                for (c, b, e) in rv.v:
                    row += "\t%s:%d:%d:-" % (c, b, e - b + 1)
                umi.append(row + "\n")
    var = ["#SNP_ID\tCHROM\tPOS\tREF\tALT\tAF\n"]
    plp = ["#DROPLET_ID\tSNP_ID\tALLELES\tBASEQS\n"]
    for i, v in enumerate(snps):
        if not demuxlet_mode:
            var.append("%d\t%s\t%d\t%s\t%s\t%.5f\n" % (i, rchroms.get(v["rid"], ""), v["pos"], v["ref"], v["alt"], v["af"]))
        for c in sorted(snp_umis.get(i, {})):
            reads = [snp_umis[i][c][u] for u in sorted(snp_umis[i][c])]
            plp.append("%d\t%d\t%s\t%s\n" % (c, i, "".join(chr(48 + a) for a, _ in reads), "".join(chr(33 + q) for _, q in reads)))
    return dict(cel="".join(cel), var="".join(var), plp="".join(plp), umi=None if umi is None else "".join(umi),
                barcodes=bcs, snps=snps, snp_umis=snp_umis, totl=totl, uniq=uniq)

