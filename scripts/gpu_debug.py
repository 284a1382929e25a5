"""First-light script for the GPU box: runs every golden case through the CUDA path and prints diffs."""
import os, sys, time, traceback
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', 'tests')))
import numpy as np, torch
from oracle.cases import GOLDEN_CASES, GOLDEN_DIR, case_inputs, case_supports
from gpu_helpers import build_model, load_synth, rel, oracle_run, compare_grads

print(torch.cuda.get_device_name(0))
for name, c in GOLDEN_CASES.items():
    try:
        cfg = c['cfg']; g = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
        sup = case_supports(c['supports'])
        m = build_model(cfg, sup, horizon=c.get('horizon', 1)); sd = load_synth(m, cfg, c['seed'])
        x_np, _ = case_inputs(name)
        x = torch.tensor(x_np, device='cuda', requires_grad=True)
        m.train()
        out = m(x)
        torch.cuda.synchronize()
        print(f'[{name}] out {tuple(out.shape)} rel_out={rel(out, g["out_train"]):.3e}', flush=True)
        loss = torch.nn.functional.mse_loss(out, torch.tensor(g['target'], device='cuda'))
        loss.backward(); torch.cuda.synchronize()
        print(f'   loss {loss.item():.6f} vs {float(g["loss"]):.6f}  x_grad rel={rel(x.grad, g["x_grad"]):.3e}')
        _, _, grads, tr = oracle_run(cfg, sd, x_np, sup, g['target'], literal=c.get('literal', False), horizon=c.get('horizon'))
        rep = []
        bad = compare_grads(m, {k: v for k, v in grads.items() if k != '__x__'}, 1e-4, rep)
        worst = sorted(rep, key=lambda t: -t[1] if t[1] == t[1] else -1e9)[:6]
        print('   worst grads:', [(k, f'{e:.2e}') for k, e in worst])
        print('   FAIL:' if bad else '   grads OK', bad[:8] if bad else '')
        for k in [k for k in g.files if k.startswith('buf/') and 'running' in k][:64]:
            e = rel(m.state_dict()[k[4:]], g[k])
            if e > 1e-5: print('   running stat mismatch', k, e)
        m.eval()
        with torch.no_grad():
            oe = m(x.detach())
        print(f'   eval rel={rel(oe, g["out_eval"]):.3e}')
    except Exception:
        traceback.print_exc()
