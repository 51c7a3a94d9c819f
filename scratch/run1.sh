cd /root/repo
python scratch/attn_split_check.py 2>&1 | tail -5
python scratch/attn_trace.py 3 16384 10 1 2>&1 | tail -8
