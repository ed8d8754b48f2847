cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 && timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "
import sys; sys.path.insert(0,'.')
import __graft_entry__ as g
g.smoke()
import numpy as np, varscot_b200 as V
from tests.util import make_case
case = make_case(seed=5, contig_lens=[30000, 45, 45, 0, 23, 9000], n_guides=7, k=8, pam='AG')
text = V.PackedText.from_ascii(case.ascii, case.offsets)
with V.ScanContext(0) as ctx:
    ctx.set_chunk_words(100)
    h,_ = ctx.scan_text(text, case.guides, 8, pam='AG', cap=8)
    h2,_ = ctx.scan(case.guides, 3)
print('hits', len(h), len(h2))
" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/memcheck.log
