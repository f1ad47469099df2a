"""How much of a C2 step is the gap BETWEEN consecutive CUDA-graph replays, and does launching the next replay on a second
stream behind an event (so that its launch processing overlaps the running step, while its kernels still start only after
the previous step's last kernel) hide it?  python scripts/gap_probe.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import stil_tta_b200 as S  # noqa: E402
from stil_tta_b200 import synth  # noqa: E402

cfg = synth.CONFIGS["C2"]()
dev = torch.device("cuda")
heads = [S.STiLHead(cfg, device=dev) for _ in range(2)]
for i, h in enumerate(heads):
    h.load(synth.make_batch(cfg, seed=2022 + i))
    h.capture()
torch.cuda.synchronize()


def timed(fn, n=3000):
    """us per call of fn, timed on the device: the side streams are joined into the current stream around the region"""
    c = torch.cuda.current_stream(dev)
    for i in range(300):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(c)
    for s_ in _side:
        s_.wait_event(e0)
    for i in range(n):
        fn(i)
    for s_ in _side:
        c.wait_stream(s_)
    e1.record(c)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


_side = []


print(f"one stream, alternating heads        : {timed(lambda i: heads[i & 1].run()):6.2f} us/step")
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
_side.extend(streams)
evs = [torch.cuda.Event(), torch.cuda.Event()]
cur = torch.cuda.current_stream(dev)
for e in evs:
    e.record(cur)


def two_streams(i):
    j = i & 1
    s = streams[j]
    s.wait_event(evs[1 - j])            # step i starts only after step i-1 has completely finished
    with torch.cuda.stream(s):
        heads[j].run()
        evs[j].record(s)


t = timed(two_streams)
cur.wait_event(evs[0]); cur.wait_event(evs[1])
print(f"two streams chained by events        : {t:6.2f} us/step")

# --- is the chained version really serial?  (1) the same head on both streams: a replay overwrites what the previous one wrote;
# (2) a 200 us spin kernel in front of every replay: serial steps must then take >= 200 us each
def two_streams_same_head(i):
    j = i & 1
    s = streams[j]
    s.wait_event(evs[1 - j])
    with torch.cuda.stream(s):
        heads[0].run()
        evs[j].record(s)


t = timed(two_streams_same_head)
cur.wait_event(evs[0]); cur.wait_event(evs[1])
print(f"two streams, SAME head               : {t:6.2f} us/step")


def two_streams_sleep(i):
    j = i & 1
    s = streams[j]
    s.wait_event(evs[1 - j])
    with torch.cuda.stream(s):
        torch.cuda._sleep(400_000)       # ~200 us at 1.9 GHz
        heads[j].run()
        evs[j].record(s)


t = timed(two_streams_sleep, 300)
cur.wait_event(evs[0]); cur.wait_event(evs[1])
print(f"two streams + 200 us spin per step   : {t:6.2f} us/step (serial if >= 200)")
t = timed(lambda i: (torch.cuda._sleep(400_000), heads[i & 1].run()), 300)
print(f"one stream  + 200 us spin per step   : {t:6.2f} us/step")
