"""Chromosome-scale driver pieces around the decoder: cut a sequence into overlapping chunks, join the predictions of the chunks.

Host-side mirror of how the reference runs whole genomes: `scripts/createAugustusJoblist.pl` plans one
`augustus --predictionStart=a --predictionEnd=b` run per chunk, the outputs are concatenated in order and `scripts/join_aug_pred.pl`
drops one version of every gene that two neighbouring chunks predicted in their overlap and renumbers the rest.  Both are restated
here with the reference's rules (same chunk borders, same break point, byte-identical joined text: tests/test_chromosome.py compares
with outputs of the two Perl scripts kept under tests/golden/), so that a batched driver can plan the windows of a chromosome, decode
them in one `augb200_decode_batch` call and join in the same process.
"""
import os
import re
import subprocess
import sys

RUN_SEPARATOR = "# This output was generated with AUGUSTUS"


def plan_chunks(start, end, chunksize, overlap=None, padding=0):
    """Chunk borders (1-based, inclusive) of one sequence: [(predictionStart, predictionEnd), ...].

    createAugustusJoblist.pl:127-152: the first chunk starts at `start - padding`; a chunk covers `chunksize` bases, clipped to
    `end + padding`; the next one starts `overlap` bases before the unclipped end of the previous one; planning stops with the first
    chunk that reaches the end.  Without --overlap the script takes 500 000 for chunks above 3 Mbp (:119-121; for smaller chunks it
    dies on a dereference, so an overlap is required there)."""
    if chunksize is None or chunksize <= 0:
        raise ValueError("Need to specify chunksize.")
    if not overlap:
        if chunksize > 3000000:
            overlap = 500000
        else:
            raise ValueError("chunksize <= 3000000 needs an explicit overlap (the reference script fails without one)")
    if overlap >= chunksize:
        raise ValueError("overlap must be smaller than chunksize (the reference script would not terminate)")
    start -= padding
    end += padding
    chunks = []
    pred_start, pred_end = start, -1
    while pred_start <= end and pred_end < end:
        pred_end = pred_start + chunksize - 1
        chunks.append((pred_start, min(pred_end, end)))
        pred_start = pred_end + 1 - overlap
    return chunks


def chunk_names(seqnr, name, chunks, outputdir):
    """Output file names of the chunk runs, `<outputdir>/<seqnr>.<chunknr %03d>.<name>.<start>..<end>.gff` (createAugustusJoblist.pl:140)."""
    name = name.rsplit("/", 1)[-1]
    return ["%s/%d.%03d.%s.%d..%d.gff" % (outputdir.rstrip("/"), seqnr, i + 1, name, a, b) for i, (a, b) in enumerate(chunks)]


# ---------------------------------------------------------------------------------------------------------------------------------
# join_aug_pred.pl

_GENE_START = re.compile(r"^### gene |^# start gene ")
_GENE_END = re.compile(r"^### end gene |^# end gene ")
_GENE_ID = re.compile(r"gene \S*(g\d+)", re.ASCII)
_GENE_LINE = re.compile(r"^(\S+)\t.*\tgene\t(\d+)\t(\d+)\t", re.ASCII)
_HEADER_END = re.compile(r"prediction on sequence number|Looks like .* is in .* format")
_RENUMBER = re.compile(r"\bg(\d+)\b", re.ASCII)


class _Gene:
    __slots__ = ("begin", "end", "gff")

    def __init__(self):
        self.begin, self.end, self.gff = 0, 0, ""       # a gene block without a `gene` feature line compares like 0 (Perl undef)


class _Run:
    __slots__ = ("header", "genes", "seqname")

    def __init__(self):
        self.header, self.genes, self.seqname = "", None, None


class Joiner:
    """join_aug_pred.pl:68-279 as a streaming filter: feed the concatenated outputs line by line (with line ends), collect `out`.

    A run is what follows one `# This output was generated with AUGUSTUS` line; its header is everything before the first gene up to
    the `prediction on sequence number` line (only the first run's header is printed); of two neighbouring runs on the same sequence
    whose genes overlap, the genes of the left run that end right of a break point and the genes of the right run that begin at or
    left of it are removed (`clean_redundant`); the remaining genes are renumbered g1, g2, ... in output order."""

    def __init__(self, droplist=None, err=None):
        self.drop = set(re.sub(r"\.t\d+", "", g.rstrip("\n"), count=1) for g in (droplist or []))
        self.err = err if err is not None else sys.stderr
        self.out = []
        self.geneid = 1
        self.header_printed = False
        self.gff3 = False

    # -- getNextRun (:87-142): the separator line that closed a run is the first line of the next one
    def _runs(self, lines):
        it = iter(lines)
        line = None
        while True:
            run, seps, gene_nr, in_gene, end_header = _Run(), 0, 0, False, False
            while True:
                if line is not None:
                    if "##gff-version 3" in line:
                        self.gff3 = True
                    if line.startswith(RUN_SEPARATOR):
                        seps += 1
                    if seps == 1:
                        if _GENE_START.search(line):
                            m = _GENE_ID.search(line)
                            if m and m.group(1) in self.drop:
                                self.err.write("dropping %s\n" % m.group(1))
                            else:
                                gene_nr += 1
                                in_gene = True
                        if gene_nr == 0:
                            if _HEADER_END.search(line):
                                end_header = True
                            if not end_header:
                                run.header += line
                        if gene_nr > 0 and in_gene:
                            if run.genes is None:
                                run.genes = []
                            while len(run.genes) < gene_nr:
                                run.genes.append(_Gene())
                            g = run.genes[gene_nr - 1]
                            m = _GENE_LINE.match(line)
                            if m:
                                if run.seqname is None:
                                    run.seqname = m.group(1)
                                g.begin, g.end = int(m.group(2)), int(m.group(3))
                            g.gff += line
                        if _GENE_END.search(line):
                            in_gene = False
                if seps < 2:
                    line = next(it, None)
                if line is None or seps > 1:
                    break
            if line is not None or gene_nr > 0:
                yield run
            else:
                return

    # -- printRun (:145-163)
    def _print_run(self, run):
        if not self.header_printed:
            self.out.append(run.header)
            self.header_printed = True
        for g in run.genes or []:
            rows = g.gff.split("\n")
            while rows and rows[-1] == "":
                rows.pop()
            for row in rows:
                self.out.append(_RENUMBER.sub("g%d" % self.geneid, row) + "\n")
            self.geneid += 1

    # -- cleanRedundant (:169-279); run1 lies left of run2
    def _clean_redundant(self, run1, run2):
        if run1.seqname is None or run2.seqname is None or run1.seqname != run2.seqname:
            return
        if not run1.genes or not run2.genes:
            return
        first_begin2 = run2.genes[0].begin
        last_end1 = -1
        for g in run1.genes:
            if g.end > last_end1:
                last_end1 = g.end
        last_end2 = -1
        for g in run2.genes:
            if last_end2 == -1 or g.end > last_end2:
                last_end2 = g.end
        if last_end2 < run1.genes[0].begin:
            self.err.write("Prediction runs are not in the right order in sequence %s. Please sort along the chromosome first.\n" % run1.seqname)
        if last_end1 < first_begin2:
            return                                      # no gene of run1 reaches the first gene of run2
        d1 = d2 = -1
        for g in run2.genes:                            # how far genes of run2 that begin inside run1's genes stick out to the right
            if g.begin <= last_end1 and g.end - last_end1 > d1:
                d1 = g.end - last_end1
        for g in run1.genes:                            # how far genes of run1 that reach run2's first gene stick out to the left
            if g.end >= first_begin2 and first_begin2 - g.begin > d2:
                d2 = first_begin2 - g.begin
        if d1 >= d2:
            breakpoint_ = last_end1                     # directly left of the leftmost gene of run2 that ends at or after last_end1
            for g in run2.genes:
                if g.end >= last_end1 and g.begin - 1 < breakpoint_:
                    breakpoint_ = g.begin - 1
        else:
            breakpoint_ = first_begin2                  # rightmost end of the genes of run1 that begin at or before first_begin2
            for g in run1.genes:
                if g.begin <= first_begin2 and g.end > breakpoint_:
                    breakpoint_ = g.end
        run1.genes = [g for g in run1.genes if g.end <= breakpoint_]
        run2.genes = [g for g in run2.genes if g.begin > breakpoint_]

    def join(self, lines):
        last = None
        for cur in self._runs(lines):
            if last is not None:
                self._clean_redundant(last, cur)
                if self.gff3 and self.geneid == 1:
                    self.out.append("##gff-version 3\n")
                self._print_run(last)
            last = cur
        if last is not None:
            self._print_run(last)
        return "".join(self.out)


def join_predictions(text, droplist=None, err=None):
    """Joined, renumbered predictions of the concatenated chunk outputs `text` (a str, or an iterable of lines with their line ends)."""
    lines = text.splitlines(keepends=True) if isinstance(text, str) else text
    return Joiner(droplist, err).join(lines)


def run_chunks(exe, fasta, chunks, args=(), env=None, timeout=3600, jobs=1):
    """One front-end process per chunk as the joblist of the reference does (`<command> --predictionStart=a --predictionEnd=b
    <fasta>`, createAugustusJoblist.pl:176-178), `jobs` of them at a time (they share the GPU); returns the outputs concatenated in
    chunk order.  `exe` is a binary with the `augustus` command line — the drop-in front end of INTEGRATION.md §0 decodes on the GPU;
    a failing run raises (no fallback)."""
    def one(chunk):
        a, b = chunk
        r = subprocess.run([exe] + list(args) + ["--predictionStart=%d" % a, "--predictionEnd=%d" % b, fasta],
                           env=env if env is not None else os.environ, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError("chunk %d..%d failed (%d): %s" % (a, b, r.returncode, (r.stderr or r.stdout)[-2000:]))
        return r.stdout
    if jobs <= 1:
        return "".join(one(c) for c in chunks)
    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(jobs) as ex:
        return "".join(ex.map(one, chunks))


def predict_chromosome(exe, fasta, length, chunksize, overlap, args=(), env=None, padding=0, jobs=1):
    """Plan the chunks of a sequence of `length` bases, run them, join: the joined GFF text of the whole sequence."""
    return join_predictions(run_chunks(exe, fasta, plan_chunks(1, length, chunksize, overlap, padding), args, env, jobs=jobs))


if __name__ == "__main__":          # filter like the reference script: chromosome.py [--droplist=file] < augustus.concat > augustus.joined
    drop = None
    for a in sys.argv[1:]:
        if a.startswith("--droplist="):
            with open(a.split("=", 1)[1]) as f:
                drop = f.readlines()
        else:
            sys.exit("Unknown option")
    sys.stdout.write(join_predictions(sys.stdin, drop))
