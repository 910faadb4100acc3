"""Debug tool: the flag barrier run by several "ranks" on several streams of one device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpqe_b200 import ops
DEV = 'cuda:0'
for world in (2, 4):
    flags = [torch.zeros(16, dtype=torch.int32, device=DEV) for _ in range(world)]
    epochs = [torch.zeros(1, dtype=torch.int32, device=DEV) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for rounds in range(3):
        for r, st in enumerate(streams):
            with torch.cuda.stream(st):
                if r == 0 and rounds == 1:
                    torch.cuda._sleep(2000000)
                ops.peer_barrier([f.data_ptr() for f in flags], r, epochs[r])
        try:
            torch.cuda.synchronize()
            print('world', world, 'round', rounds, 'ok; epochs', [int(e) for e in epochs], 'flags', [f[:world].tolist() for f in flags])
        except Exception as exc:
            print('world', world, 'round', rounds, 'FAILED', repr(exc)[:200])
            sys.exit(1)
