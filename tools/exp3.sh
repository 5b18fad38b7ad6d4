set -x
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'.')
from tests import util
from augustus_b200 import synth
name,dna=util.read_fasta(util.GOLDEN+'/fly_softmask_window.fa')[0]
synth.write_fasta('/tmp/fly.fa',[dna],[name])
PY
AUGUSTUS_CONFIG_PATH=$PWD/oracle/_ref/config AUGSHIM_VERBOSE=1 oracle/_ref/augustus_b200 --species=fly /tmp/fly.fa > /tmp/fly_b200.gff 2> /tmp/fly_b200.err; echo rc=$?
tail -5 /tmp/fly_b200.err; tail -5 /tmp/fly_b200.gff
AUGUSTUS_CONFIG_PATH=$PWD/oracle/_ref/config oracle/_ref/augustus --species=fly /tmp/fly.fa > /tmp/fly_ref.gff 2>/tmp/fly_ref.err; echo rc=$?
diff /tmp/fly_ref.gff /tmp/fly_b200.gff | head -20
python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/prof_sweep.py 2368 3 2>&1 | tail -1
