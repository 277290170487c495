#!/usr/bin/env python
"""TF checkpoint (V2 bundle) of an EPOS model -> .npz of TF-named variables for epos_b200 (no TensorFlow needed).

  python scripts/convert_checkpoint.py /path/to/model.ckpt-123456 weights.npz [--model_variant xception_65]

The reference restores the same files with tf.train.Saver (/root/reference/scripts/infer.py:670-683)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('checkpoint', help='checkpoint prefix (the path without .index / .data-*)')
    ap.add_argument('output', help='.npz to write')
    ap.add_argument('--model_variant', default='xception_65', choices=['xception_65', 'resnet_v1_50_beta'])
    ap.add_argument('--list', action='store_true', help='only list the variables of the checkpoint')
    args = ap.parse_args()
    from epos_b200 import tf_checkpoint, weights as W
    if args.list:
        for name, shape, dt in tf_checkpoint.list_variables(args.checkpoint):
            print('%-90s %-22s %s' % (name, shape, getattr(dt, '__name__', dt)))
        return
    w, O, F = tf_checkpoint.epos_weights_from_checkpoint(args.checkpoint, args.model_variant)
    W.save_npz(args.output, w)
    print('wrote %d variables (%d objects x %d fragments, %s) to %s' % (len(w), O, F, args.model_variant, args.output))


if __name__ == '__main__':
    main()
