"""Read / ground-truth / database preparation for the alignment drivers — dependency-free (no pandas).

The reference prepares its inputs with py/reader.py and py/ompfg_data_prep.py (pandas `DataFrame.append`, which current
pandas no longer has).  This module writes the SAME files, so that the drivers of both trees read them:

  sam_to_ground_truth   SAM -> "index,QNAME,SEQ,POS" CSV          py/reader.py:27-43 (class SAM), :161-173 (gen_input_125)
  fastq_sequences       every 4th line of a FASTQ (line 1, 5, ...) py/reader.py:45-50 (mpi_prepare), :107-115
  read_fa / single_line_fa  header dropped, lines concatenated     py/reader.py:117-123, :141-159
  gen_ref_custom        a window of a long FASTA, upper-cased, N removed   py/ompfg_data_prep.py:31-73
  gen_reads_custom      reads sampled uniformly from a reference + ground-truth CSV + "_readsonly.txt"   :75-122
  split_multifasta      '>sp' records -> one sequence per line     py/reader.py:75-97 (uniprot_prepare_single)
  pack_database / load_database   multi-FASTA -> ONE packed blob (5-bit residue codes) + offsets, sorted by length:
                        replaces the 561 356 one-protein FASTA files of py/reader.py:52-73 that
                        mpi_sw_solve_uniprot.cpp:97-110 opens one by one

Command line:  python -m parallel-genomeseq_b200.dataprep <command> ...   (see main()).
"""
import os
import random
import struct
import sys

import numpy as np

SAM_FIELDS = ["QNAME", "FLAG", "RNAME", "POS", "MAPQ", "CIGAR", "RNEXT", "PNEXT", "TLEN", "SEQ", "QUAL"]


# ---- SAM / FASTQ / FASTA ------------------------------------------------------------------------------------
def parse_sam(text):
    """py/reader.py:27-43: '@' lines are metadata, every other line is tab-separated into the eleven SAM fields (further
    optional fields are ignored).  Returns (meta_lines, rows) with rows as dicts."""
    meta, rows = [], []
    for line in text.split("\n"):
        if not line:
            continue
        if line[0] == "@":
            meta.append(line)
            continue
        f = line.split("\t")
        rows.append({SAM_FIELDS[j]: f[j] for j in range(min(len(f), len(SAM_FIELDS)))})
    return meta, rows


def sam_to_ground_truth(sam_path, csv_path):
    """py/reader.py:161-173 (gen_input_125): the ground-truth CSV the drivers read — header "index,QNAME,SEQ,POS", one row
    per SAM record, index counting from 0 (sw_solve_small.cpp:56-67 takes field 2 as the read)."""
    with open(sam_path) as f:
        _, rows = parse_sam(f.read())
    with open(csv_path, "w") as f:
        f.write("index,QNAME,SEQ,POS\n")
        for i, r in enumerate(rows):
            f.write(f"{i},{r['QNAME']},{r['SEQ']},{r['POS']}\n")
    return len(rows)


def fastq_sequences(fq_path, out_path=None):
    """py/reader.py:45-50: the sequence line of every FASTQ record (lines 1, 5, 9, ... counting from 0), one per line —
    the 126-byte-per-line file mpi_sw_solve_small.cpp:57-76 reads with MPI-IO."""
    with open(fq_path) as f:
        text = f.read().split("\n")
    seqs = [text[i] for i in range(1, len(text), 4)]
    if out_path:
        with open(out_path, "w") as f:
            for s in seqs:
                f.write(s + "\n")
    return seqs


def read_fa(path):
    """py/reader.py:117-123 and sw_solve_small.cpp:25-30: skip the header line, concatenate the rest."""
    with open(path) as f:
        return "".join(f.read().split("\n")[1:])


def single_line_fa(in_path, out_path):
    """py/reader.py:141-159: the whole reference as one header-less line (what sw_solve_big.cpp:36 reads)."""
    with open(out_path, "w") as out, open(in_path) as f:
        for i, line in enumerate(f):
            if i > 0:
                out.write(line.rstrip("\n"))


def gen_ref_custom(long_fa, out_path, start_pos=300000 * 60, ref_len=30 * 1000, remove_n=True):
    """py/ompfg_data_prep.py:31-73: the bases [start_pos, start_pos + ref_len) of a long FASTA (whole lines that START in
    the window, as the reference does), upper-cased, 'N' removed, written as one header-less line."""
    nt, out = 0, []
    with open(long_fa) as f:
        for i, line in enumerate(f):
            if i == 0:
                continue
            s = line.rstrip("\n").upper()
            if start_pos <= nt < start_pos + ref_len:
                out.append(s)
            nt += len(s)
    ref = "".join(out)
    if remove_n:
        ref = ref.replace("N", "")
    with open(out_path, "w") as f:
        f.write(ref)
    return ref


def gen_reads_custom(ref_path, out_csv, read_len=10000, n_reads=100, seed=None):
    """py/ompfg_data_prep.py:75-122: n_reads windows of read_len bases at uniform positions start = int(random() * (len(ref) -
    read_len)); ground-truth CSV "index,QNAME,SEQ,POS" (POS = 0-based start) plus "<stem>_readsonly.txt"."""
    with open(ref_path) as f:
        ref = f.read()
    rng = random.Random(seed)
    rows = []
    for i in range(n_reads):
        start = int(rng.random() * (len(ref) - read_len))
        rows.append((f"custom_read_{i}", ref[start:start + read_len], start))
    with open(out_csv, "w") as f:
        f.write("index,QNAME,SEQ,POS\n")
        for i, (q, s, p) in enumerate(rows):
            f.write(f"{i},{q},{s},{p}\n")
    with open(out_csv.split(".")[0] + "_readsonly.txt", "w") as f:
        for _, s, _ in rows:
            f.write(s + "\n")
    return rows


def fasta_records(path, token=">"):
    """Multi-FASTA records as (header, sequence); py/reader.py:52-97 splits UniProt on the '>sp' token."""
    recs, head, cur = [], None, []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(token):
                if head is not None or cur:
                    recs.append((head or "", "".join(cur)))
                head, cur = line, []
            elif not line.startswith(">"):
                cur.append(line)
    if head is not None or cur:
        recs.append((head or "", "".join(cur)))
    return recs


def split_multifasta(path, out_path, stats_path=None):
    """py/reader.py:75-97 (uniprot_prepare_single): one sequence per line in database.fasta, the count in stats.txt."""
    recs = fasta_records(path)
    with open(out_path, "w") as f:
        for _, s in recs:
            f.write(s + "\n")
    if stats_path:
        with open(stats_path, "w") as f:
            f.write("%d" % len(recs))
    return len(recs)


# ---- packed protein / nucleotide database --------------------------------------------------------
MAGIC = b"SWBDB001"
ALPHABET = b"ARNDCQEGHILKMFPSTWYVBZX*UOJ-"      # code = index; anything else -> 'X'; 28 symbols fit 5 bits


def pack_database(fasta_path, out_path):
    """Multi-FASTA -> one file: header, offsets and the residues as 5-bit codes (8 residues in 5 bytes), entries sorted by
    DECREASING length (the order the kernels want: pair-mates of similar length, long work first); the original index of
    every entry is kept so results can be written in input order.

    Layout (little endian): MAGIC[8] | n_entries u64 | n_residues u64 | alphabet_len u32 | alphabet bytes (padded to 4) |
    orig_index u32[n] | offsets u64[n+1] (residue units, in sorted order) | packed codes ceil(n_residues * 5 / 8) bytes."""
    recs = fasta_records(fasta_path)
    lut = np.full(256, ALPHABET.index(b"X"), np.uint8)
    for i, c in enumerate(ALPHABET):
        lut[c] = i
        lut[ord(chr(c).lower())] = i
    order = sorted(range(len(recs)), key=lambda i: -len(recs[i][1]))
    lens = np.array([len(recs[i][1]) for i in order], dtype=np.uint64)
    offs = np.zeros(len(order) + 1, np.uint64)
    offs[1:] = np.cumsum(lens)
    blob = np.frombuffer("".join(recs[i][1] for i in order).encode("latin-1"), dtype=np.uint8)
    codes = lut[blob]
    n = len(codes)
    pad = (-n) % 8
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]).astype(np.uint64).reshape(-1, 8)
    word = np.zeros(len(c), np.uint64)
    for k in range(8):
        word |= c[:, k] << np.uint64(5 * k)
    packed = np.zeros((len(c), 5), np.uint8)
    for b in range(5):
        packed[:, b] = (word >> np.uint64(8 * b)) & np.uint64(0xFF)
    alpha = ALPHABET + b"\0" * ((-len(ALPHABET)) % 4)
    with open(out_path, "wb") as f:
        f.write(MAGIC + struct.pack("<QQI", len(order), n, len(ALPHABET)) + alpha)
        f.write(np.asarray(order, np.uint32).tobytes())
        f.write(offs.tobytes())
        f.write(packed.tobytes())
    return len(order), n


def load_database(path):
    """-> (blob uint8[n_residues] of residue BYTES, offsets uint64[n+1], orig_index uint32[n]) in the stored (length-sorted)
    order; (blob, offsets) is exactly what Engine.stage / swb_batch_stage take."""
    with open(path, "rb") as f:
        raw = f.read()
    assert raw[:8] == MAGIC, "not a packed database"
    n_ent, n_res, alen = struct.unpack_from("<QQI", raw, 8)
    p = 8 + 20
    alpha = np.frombuffer(raw, np.uint8, alen, p)
    p += alen + ((-alen) % 4)
    orig = np.frombuffer(raw, np.uint32, n_ent, p)
    p += 4 * n_ent
    offs = np.frombuffer(raw, np.uint64, n_ent + 1, p)
    p += 8 * (n_ent + 1)
    nw = (n_res + 7) // 8
    packed = np.frombuffer(raw, np.uint8, nw * 5, p).reshape(-1, 5).astype(np.uint64)
    word = np.zeros(nw, np.uint64)
    for b in range(5):
        word |= packed[:, b] << np.uint64(8 * b)
    codes = np.zeros((nw, 8), np.uint8)
    for k in range(8):
        codes[:, k] = (word >> np.uint64(5 * k)) & np.uint64(31)
    blob = alpha[codes.reshape(-1)[:n_res]]
    return np.ascontiguousarray(blob), offs.copy(), orig.copy()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    cmd, a = argv[0], argv[1:]
    if cmd == "sam2csv":
        print(sam_to_ground_truth(a[0], a[1]), "reads")
    elif cmd == "fastq2lines":
        print(len(fastq_sequences(a[0], a[1])), "reads")
    elif cmd == "single_line_fa":
        single_line_fa(a[0], a[1])
    elif cmd == "gen_ref_custom":
        print(len(gen_ref_custom(a[0], a[1], int(a[2]) if len(a) > 2 else 300000 * 60, int(a[3]) if len(a) > 3 else 30000)), "bases")
    elif cmd == "gen_reads_custom":
        print(len(gen_reads_custom(a[0], a[1], int(a[2]) if len(a) > 2 else 10000, int(a[3]) if len(a) > 3 else 100, int(a[4]) if len(a) > 4 else None)), "reads")
    elif cmd == "split_multifasta":
        print(split_multifasta(a[0], a[1], a[2] if len(a) > 2 else None), "records")
    elif cmd == "pack_db":
        print("%d entries, %d residues" % pack_database(a[0], a[1]))
    else:
        print(__doc__)
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main())
