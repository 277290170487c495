"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle (test infrastructure;
this module and tests/ and bench.py's cpu_baseline leg are the only importers of oracle/)."""
import numpy as np
import torch


def run():
    from . import model, weights as W
    from oracle import cnn as ocnn
    assert torch.cuda.is_available(), 'smoke() needs a CUDA device'
    dev = torch.device('cuda:0')
    O, F = 2, 8
    w = W.random_init(O, F, seed=3, bn='random', logits_std=0.5)
    img = W.synthetic_images(1, seed=3, height=96, width=128)
    net = model.EposNet(w, O, F, dev)
    out = net.predict(torch.from_numpy(img).to(dev))
    torch.cuda.synchronize()
    ref = ocnn.predict(w, img, O, F)
    for k in (model.PRED_OBJ_CONF, model.PRED_FRAG_CONF, model.PRED_FRAG_LOC):
        a, b = out[k].cpu().numpy(), ref[k]
        err = np.abs(a - b).max() / np.abs(b).max()
        print('[smoke] %-15s rel-to-max err %.3e' % (k, err))
        assert err < 1e-3, (k, err)
    agree = (out[model.PRED_OBJ_LABEL].cpu().numpy() == ref['pred_obj_label']).mean()
    print('[smoke] label agreement %.5f' % agree)
    assert agree > 0.999
    print('[smoke] OK')
