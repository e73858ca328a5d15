"""Synthetic "simuxlet-style" demuxlet / freemuxlet inputs.

Random genotypes, simulated singlet / doublet droplets, reads drawn from the droplet's sample(s),
emitted directly as the float GP matrix plus the CSR droplet store that
`sc_dropseq_lib_t::load_from_plp` (sc_drop_seq.cpp:103-384) would build.  Counter-based PRNG
(numpy Philox), seed = 1000 + config index.  Reads per cell r-bar is OUR assumption for C2..C4
(BASELINE.json does not fix it) and is recorded in every config dict.
"""
import numpy as np

# name -> BASELINE.json configs[i]
CONFIGS = {
    "C1": dict(index=0, nv=8, nsnps=20_000, nbcs=1_000, rbar=50, alphas=[0.0, 0.5], doublet_rate=0.10),
    "C2": dict(index=1, nv=64, nsnps=200_000, nbcs=20_000, rbar=200,
               alphas=[round(0.05 * i, 2) for i in range(11)], doublet_rate=0.10),
    "C3": dict(index=2, nv=256, nsnps=500_000, nbcs=100_000, rbar=200,
               alphas=[round(0.05 * i, 2) for i in range(11)], doublet_rate=0.10),
    "C4": dict(index=3, nv=16, nsnps=300_000, nbcs=50_000, rbar=200, alphas=[0.0, 0.5], doublet_rate=0.10),
    "C5": dict(index=4, nv=128, nsnps=200_000, nbcs=5_000, rbar=2000,
               alphas=[round(0.05 * i, 2) for i in range(11)], doublet_rate=0.20),
}


def _hex_order_key(counter):
    """Sort key reproducing lexicographic order of sprintf("%x", counter) (sc_drop_seq.cpp:365:
    UMIs in the plp path are hex strings of a global counter kept in a std::map<std::string,..>)."""
    c = counter.astype(np.uint64)
    nd = np.ones(c.shape, np.int64)
    t = c >> np.uint64(4)
    while np.any(t > 0):
        nd += (t > 0)
        t = t >> np.uint64(4)
    left = c << ((16 - nd) * 4).astype(np.uint64)
    return left, nd


def make_genotypes(rng, nsnps, nv):
    """float32 GP [nsnps][nv][3], normalised in float like bcf_filtered_reader.cpp:474-485."""
    af = rng.uniform(0.05, 0.5, nsnps)
    geno = rng.binomial(2, af[:, None], size=(nsnps, nv)).astype(np.int8)
    gp = np.full((nsnps, nv, 3), np.float32(0.01), np.float32)
    np.put_along_axis(gp, geno[:, :, None].astype(np.int64), np.float32(0.98), axis=2)  # the VCF's GP values
    s = (gp[:, :, 0] + gp[:, :, 1]) + gp[:, :, 2]  # float accumulation order of :478-480
    gp /= s[:, :, None]
    return af, geno, gp


def make_droplets(rng, geno, nbcs, rbar, doublet_rate, min_bq=13, cap_bq=20, chunk=2_000_000, snp_skew=0.0):
    """Simulated droplets -> CSR (cell_ptr, snp_idx, read_ptr, al, bq) + truth.
    snp_skew > 0: the SNP of a read is drawn from a Zipf law P(rank r) ~ 1/r^snp_skew over a random permutation of
    the panel instead of uniformly -- expression in real scRNA-seq is skewed, so a few SNPs collect many reads of a
    droplet and a large share of the (cell, SNP) entries holds two or more reads."""
    nsnps, nv = geno.shape
    is_dbl = rng.random(nbcs) < doublet_rate
    s1 = rng.integers(0, nv, nbcs)
    s2 = (s1 + 1 + rng.integers(0, max(nv - 1, 1), nbcs)) % nv
    s2 = np.where(is_dbl, s2, s1)
    nreads = rng.poisson(rbar, nbcs).astype(np.int64)
    R = int(nreads.sum())
    cell = np.repeat(np.arange(nbcs, dtype=np.int64), nreads)
    if snp_skew > 0:
        w = 1.0 / np.power(np.arange(1, nsnps + 1, dtype=np.float64), snp_skew)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        perm = rng.permutation(nsnps)
        snp = perm[np.minimum(np.searchsorted(cdf, rng.random(R), side="right"), nsnps - 1)].astype(np.int64)
    else:
        snp = rng.integers(0, nsnps, R).astype(np.int64)
    src = np.where(rng.random(R) < 0.5, s1[cell], s2[cell])
    g = np.empty(R, np.int8)
    for a in range(0, R, chunk):  # gather in chunks to bound temporaries
        b = min(R, a + chunk)
        g[a:b] = geno[snp[a:b], src[a:b]]
    raw_bq = rng.integers(min_bq, 41, R)
    allele = (rng.random(R) < g * 0.5).astype(np.uint8)
    flip = rng.random(R) < np.power(10.0, -raw_bq / 10.0)
    allele = np.where(flip, 1 - allele, allele).astype(np.uint8)
    allele = np.where(rng.random(R) < 0.001, 2, allele).astype(np.uint8)
    bq = np.minimum(raw_bq, cap_bq).astype(np.uint8)

    # plp order: by droplet, then SNP; the global hex counter follows that order
    order = np.lexsort((np.arange(R), snp, cell))
    cell, snp, allele, bq = cell[order], snp[order], allele[order], bq[order]
    counter = np.arange(R, dtype=np.int64)
    newgrp = np.ones(R, bool)
    if R > 1:
        newgrp[1:] = (cell[1:] != cell[:-1]) | (snp[1:] != snp[:-1])
    grp = np.cumsum(newgrp) - 1
    nnz = int(grp[-1] + 1) if R else 0
    gsize = np.bincount(grp, minlength=nnz)
    multi = gsize[grp] > 1
    if np.any(multi):  # reads of one (cell,SNP) iterate in lexicographic UMI order
        idx = np.nonzero(multi)[0]
        left, nd = _hex_order_key(counter[idx])
        sub = np.lexsort((nd, left, grp[idx]))
        allele[idx] = allele[idx][sub]
        bq[idx] = bq[idx][sub]
    read_ptr = np.zeros(nnz + 1, np.int64)
    np.cumsum(gsize, out=read_ptr[1:])
    first = np.nonzero(newgrp)[0]
    snp_idx = snp[first].astype(np.int32)
    ent_cell = cell[first]
    cell_ptr = np.zeros(nbcs + 1, np.int64)
    np.cumsum(np.bincount(ent_cell, minlength=nbcs), out=cell_ptr[1:])
    return dict(cell_ptr=cell_ptr, snp_idx=snp_idx, read_ptr=read_ptr, al=allele, bq=bq,
                is_dbl=is_dbl, s1=s1, s2=s2, nreads=nreads)


def make_barcodes(rng, nbcs):
    out, seen = [], set()
    letters = np.array(list("ACGT"))
    while len(out) < nbcs:
        for row in letters[rng.integers(0, 4, (nbcs, 16))]:
            b = "".join(row) + "-1"
            if b not in seen:
                seen.add(b)
                out.append(b)
                if len(out) == nbcs:
                    break
    return out


def make_dataset(name="C1", nbcs=None, nsnps=None, nv=None, rbar=None, seed=None, droplet_seed_offset=0,
                 geno_error_offset=0.10, with_barcodes=False, snp_skew=0.0):
    """Dataset of a named BASELINE config (any size field can be overridden for tests)."""
    cfg = dict(CONFIGS[name])
    if nbcs is not None:
        cfg["nbcs"] = nbcs
    if nsnps is not None:
        cfg["nsnps"] = nsnps
    if nv is not None:
        cfg["nv"] = nv
    if rbar is not None:
        cfg["rbar"] = rbar
    seed = 1000 + cfg["index"] if seed is None else seed
    rng_g = np.random.Generator(np.random.Philox(key=seed))
    af, geno, gp = make_genotypes(rng_g, cfg["nsnps"], cfg["nv"])
    rng_d = np.random.Generator(np.random.Philox(key=seed + 7919 * (1 + droplet_seed_offset)))
    d = make_droplets(rng_d, geno, cfg["nbcs"], cfg["rbar"], cfg["doublet_rate"], snp_skew=snp_skew)
    d.update(snp_skew=snp_skew, name=name, cfg=cfg, seed=seed, gp=gp, af=af, geno=geno,
             snp_err=np.full(cfg["nsnps"], geno_error_offset, np.float64),
             alphas=np.asarray(cfg["alphas"], np.float64), nv=cfg["nv"], nsnps=cfg["nsnps"], nbcs=cfg["nbcs"],
             sample_ids=["SM%03d" % i for i in range(cfg["nv"])])
    if with_barcodes:
        d["barcodes"] = make_barcodes(rng_d, cfg["nbcs"])
    return d


def algorithmic_flops_per_entry(nv, alphas):
    """FP64 flops per (cell,SNP) = 8*N_llk + 18*nv*nA."""
    nA = len(alphas)
    has_half = any(a == 0.5 for a in alphas)
    n_llk = nv + nv * (nv - 1) * (nA - 1) - (nv * (nv - 1) // 2 if has_half else 0)
    return 8 * n_llk + 18 * nv * nA


# ---------------------------------------------------------------------------------------------
# on-disk forms of a dataset: digital pileup (.cel.gz/.var.gz/.plp.gz) + VCF
# ---------------------------------------------------------------------------------------------
def plp_read_order(d):
    """Index array `o` such that reads[o] is the order in which the reads were numbered by the global
    hex counter of load_from_plp (sc_drop_seq.cpp:365), i.e. the order they must appear in .plp.gz.
    The CSR keeps each entry's reads in std::map order of that counter's hex string."""
    rp = np.asarray(d["read_ptr"], np.int64)
    R = int(rp[-1])
    o = np.arange(R, dtype=np.int64)
    cnt = np.diff(rp)
    for e in np.nonzero(cnt > 1)[0]:
        c = np.arange(rp[e], rp[e + 1], dtype=np.int64)  # counters of this entry (consecutive)
        left, nd = _hex_order_key(c)
        perm = np.lexsort((nd, left))                    # CSR position j holds counter c[perm[j]]
        o[c[perm]] = c
    return o


def write_plp(d, prefix, chrom="1"):
    """Writes <prefix>.cel.gz/.var.gz/.plp.gz in the dsc-pileup format read by load_from_plp
    (sc_drop_seq.cpp:103-384; writer side: cmd_cram_dsc_pileup.cpp:438-523)."""
    import gzip
    cp, rp = np.asarray(d["cell_ptr"], np.int64), np.asarray(d["read_ptr"], np.int64)
    nbcs = len(cp) - 1
    bcs = d.get("barcodes") or ["BC%07d-1" % i for i in range(nbcs)]
    nread_cell = rp[cp[1:]] - rp[cp[:-1]]
    with gzip.open(prefix + ".cel.gz", "wt", compresslevel=1) as f:
        f.write("#DROPLET_ID\tBARCODE\tNUM.READ\tNUM.UMI\tNUM.UMIwSNP\tNUM.SNP\n")
        for i in range(nbcs):
            f.write("%d\t%s\t%d\t%d\t%d\t%d\n" % (i, bcs[i], nread_cell[i], nread_cell[i], nread_cell[i],
                                                   cp[i + 1] - cp[i]))
    with gzip.open(prefix + ".var.gz", "wt", compresslevel=1) as f:
        f.write("#SNP_ID\tCHROM\tPOS\tREF\tALT\tAF\n")
        for s in range(d["nsnps"]):
            f.write("%d\t%s\t%d\tA\tG\t%.5f\n" % (s, chrom, 1000 + 10 * s, d["af"][s]))
    order = plp_read_order(d)
    al = np.asarray(d["al"])[order]
    bq = np.asarray(d["bq"])[order]
    alc = (al + ord("0")).astype(np.uint8).tobytes().decode("ascii")
    bqc = (bq + 33).astype(np.uint8).tobytes().decode("ascii")
    si = np.asarray(d["snp_idx"])
    with gzip.open(prefix + ".plp.gz", "wt", compresslevel=1) as f:
        f.write("#DROPLET_ID\tSNP_ID\tALLELES\tBASEQS\n")
        for c in range(nbcs):
            for e in range(cp[c], cp[c + 1]):
                f.write("%d\t%d\t%s\t%s\n" % (c, si[e], alc[rp[e]:rp[e + 1]], bqc[rp[e]:rp[e + 1]]))
    return bcs


def write_vcf(d, path, chrom="1", with_r2=False, lead_decoy=False):
    """Text VCF with GT:GP per sample: the un-normalised GP triple (0.98 on the genotype, 0.01 elsewhere)
    that BCFFilteredReader::parse_posteriors (bcf_filtered_reader.cpp:457-499) float-normalises into d["gp"].
    lead_decoy: one extra heterozygous record in front of every read (position 500).  The reference's --sam path
    treats the FIRST variant of the file differently from the others (cmd_cram_demuxlet.cpp:177, 204-209: default
    genotype-error offset, no error mix); with the decoy in that role the real SNPs are handled like on --plp."""
    import gzip
    opener = gzip.open if path.endswith(".gz") else open
    gts = ["0/0", "0/1", "1/1"]
    gps = ["0.98,0.01,0.01", "0.01,0.98,0.01", "0.01,0.01,0.98"]
    cols = [gts[g] + ":" + gps[g] for g in range(3)]
    with opener(path, "wt") as f:
        f.write("##fileformat=VCFv4.2\n##contig=<ID=%s>\n" % chrom)
        f.write('##INFO=<ID=R2,Number=1,Type=Float,Description="imputation r2">\n')
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        f.write('##FORMAT=<ID=GP,Number=G,Type=Float,Description="Genotype posterior">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(d["sample_ids"]) + "\n")
        geno = np.asarray(d["geno"])
        if lead_decoy:
            f.write("%s\t500\t.\tA\tC\t.\tPASS\t%s\tGT:GP\t" % (chrom, "R2=0.9" if with_r2 else "."))
            f.write("\t".join(cols[1] for _ in d["sample_ids"]) + "\n")
        for s in range(d["nsnps"]):
            info = "R2=0.9" if with_r2 else "."
            f.write("%s\t%d\t.\tA\tG\t.\tPASS\t%s\tGT:GP\t" % (chrom, 1000 + 10 * s, info))
            f.write("\t".join(cols[g] for g in geno[s]))
            f.write("\n")


# ---------------------------------------------------------------------------------------------
# BAM writer (SAM spec section 4: BGZF blocks around little-endian records) for the --sam tests
# ---------------------------------------------------------------------------------------------
_CIGAR_OPS = "MIDNSHP=X"
_NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _bgzf_block(data):
    import struct
    import zlib
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25  # total block size - 1
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bcf(d, path, chrom="1", with_r2=False, lead_decoy=False):
    """The records of write_vcf as BCF2.2 (VCF spec section 6: BGZF around typed binary records), written from the
    specification with Python's struct -- an encoder independent of the reader under test.  GP values are the float32
    of the VCF's decimals, which is what a text -> BCF conversion stores."""
    import struct
    text = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##contig=<ID=%s>\n"
            '##INFO=<ID=R2,Number=1,Type=Float,Description="imputation r2">\n'
            '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
            '##FORMAT=<ID=GP,Number=G,Type=Float,Description="Genotype posterior">\n'
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" % chrom) + "\t".join(d["sample_ids"]) + "\n"
    PASS, R2, GT, GP = 0, 1, 2, 3  # dictionary positions: order of first appearance, PASS first
    hdr = text.encode() + b"\0"
    blob = [b"BCF\2\2", struct.pack("<I", len(hdr)), hdr]
    geno = np.asarray(d["geno"])
    nv = len(d["sample_ids"])
    gp_of = {0: (0.98, 0.01, 0.01), 1: (0.01, 0.98, 0.01), 2: (0.01, 0.01, 0.98)}

    def record(pos1, alt, genos):
        shared = struct.pack("<iiiI", 0, pos1 - 1, 1, 0x7F800001)                 # CHROM, POS0, rlen, QUAL missing
        shared += struct.pack("<II", (2 << 16) | (1 if with_r2 else 0), (2 << 24) | nv)
        shared += b"\x07"                                                        # ID: empty string
        shared += b"\x17A" + b"\x17" + alt.encode()                              # REF, ALT
        shared += b"\x11" + struct.pack("<b", PASS)                               # FILTER
        if with_r2:
            shared += b"\x11" + struct.pack("<b", R2) + b"\x15" + struct.pack("<f", 0.9)
        indiv = b"\x11" + struct.pack("<b", GT) + b"\x21"
        indiv += b"".join(struct.pack("<bb", ((1 if g >= 1 else 0) + 1) << 1, ((1 if g == 2 else 0) + 1) << 1) for g in genos)
        indiv += b"\x11" + struct.pack("<b", GP) + b"\x35"
        indiv += b"".join(struct.pack("<fff", *gp_of[int(g)]) for g in genos)
        return struct.pack("<II", len(shared), len(indiv)) + shared + indiv

    if lead_decoy:
        blob.append(record(500, "C", [1] * nv))
    for s in range(d["nsnps"]):
        blob.append(record(1000 + 10 * s, "G", geno[s]))
    data = b"".join(blob)
    with open(path, "wb") as f:
        for i in range(0, len(data), 60000):
            f.write(_bgzf_block(data[i:i + 60000]))
        f.write(_bgzf_block(b""))


def bgzip(src, dst, block=60000):
    """src re-written as BGZF (what `bgzip` / `bcftools view -Oz` produce): independent gzip members + EOF marker."""
    data = open(src, "rb").read()
    with open(dst, "wb") as f:
        for i in range(0, len(data), block):
            f.write(_bgzf_block(data[i:i + block]))
        f.write(_bgzf_block(b""))


def encode_bam_record(tid, pos, mapq, flag, cigar, seq, qual, qname="r", tags=()):
    """One alignment record.  cigar: [(op char, length)]; qual: raw Phred values; tags: [(two-letter tag, str)]."""
    import struct
    name = qname.encode() + b"\0"
    cig = b"".join(struct.pack("<I", (n << 4) | _CIGAR_OPS.index(op)) for op, n in cigar)
    packed = bytearray((len(seq) + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= _NT16[c] << (4 if i % 2 == 0 else 0)
    aux = b"".join(t.encode() + b"Z" + v.encode() + b"\0" for t, v in tags)
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), mapq, 4680, len(cigar), flag, len(seq), -1, -1, 0)
    body += name + cig + bytes(packed) + bytes(bytearray(qual)) + aux
    return struct.pack("<i", len(body)) + body


def write_bam_records(path, refs, records):
    """refs: [(name, length)]; records: encoded alignment records in coordinate order."""
    import struct
    text = b"@HD\tVN:1.6\tSO:coordinate\n" + b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (n.encode(), ln) for n, ln in refs)
    hdr = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for n, ln in refs:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln)
    blob = hdr + b"".join(records)
    with open(path, "wb") as f:
        for i in range(0, len(blob), 60000):
            f.write(_bgzf_block(blob[i:i + 60000]))
        f.write(_bgzf_block(b""))  # BGZF EOF marker


def write_bam(d, path, chrom="1", seed=0, duplicate_every=0):
    """The reads of the synthetic droplet store as a coordinate-sorted BAM with CB/UB tags: one short alignment per
    read, covering only its own SNP (SNPs of write_vcf / write_plp sit 10 bp apart at 1000 + 10*s, REF A, ALT G), with
    a CIGAR drawn from a few shapes (soft clip, insertion, deletion, reference skip) so that the read coordinate of
    the SNP base differs from its reference offset.  duplicate_every > 0 appends, after every such read, a second
    alignment with the same droplet, SNP and UMI but another base: the reference keeps the first one."""
    rng = np.random.default_rng(seed)
    cp, rp = np.asarray(d["cell_ptr"], np.int64), np.asarray(d["read_ptr"], np.int64)
    si, al, bq = np.asarray(d["snp_idx"]), np.asarray(d["al"]), np.asarray(d["bq"])
    bcs = d["barcodes"]
    base_of = {0: "A", 1: "G", 2: "C"}
    # (cigar, reference offset of the first aligned base relative to the SNP, read index of the SNP base)
    shapes = [
        ([("M", 5)], -2, 2),
        ([("S", 2), ("M", 5)], -3, 5),
        ([("M", 2), ("I", 1), ("M", 3)], -3, 4),
        ([("M", 2), ("D", 2), ("M", 3)], -5, 3),
        ([("M", 1), ("N", 3), ("M", 4), ("S", 1)], -4, 1),
        ([("H", 3), ("M", 4)], 0, 0),
    ]
    recs = []
    n = 0
    for c in range(len(cp) - 1):
        for e in range(cp[c], cp[c + 1]):
            for r in range(rp[e], rp[e + 1]):
                cigar, off, ridx = shapes[int(rng.integers(len(shapes)))]
                snp_pos0 = 1000 + 10 * int(si[e]) - 1
                rlen = sum(k for op, k in cigar if op in "MIS=X")
                seq = ["T"] * rlen
                qual = [30] * rlen
                seq[ridx] = base_of[int(al[r])]
                qual[ridx] = int(bq[r])
                umi = "%x" % n
                recs.append((snp_pos0 + off, n, encode_bam_record(0, snp_pos0 + off, 60, 0, cigar, "".join(seq), qual,
                                                                  "q%d" % n, (("CB", bcs[c]), ("UB", umi)))))
                if duplicate_every and n % duplicate_every == 0:
                    seq2 = list(seq)
                    seq2[ridx] = "G" if seq[ridx] != "G" else "A"
                    recs.append((snp_pos0 + off, n + 0.5, encode_bam_record(0, snp_pos0 + off, 60, 0, cigar, "".join(seq2),
                                                                            qual, "d%d" % n, (("CB", bcs[c]), ("UB", umi)))))
                n += 1
    recs.sort(key=lambda t: (t[0], t[1]))
    # reads the filter must drop: low mapping quality, duplicate flag, unmapped
    junk = [encode_bam_record(0, 1000, 5, 0, [("M", 5)], "GGGGG", [30] * 5, "lowmq", (("CB", bcs[0]), ("UB", "zz1"))),
            encode_bam_record(0, 1000, 60, 0x400, [("M", 5)], "GGGGG", [30] * 5, "dup", (("CB", bcs[0]), ("UB", "zz2")))]
    write_bam_records(path, [(chrom, 1000 + 10 * d["nsnps"] + 100)], junk + [t[2] for t in recs])
