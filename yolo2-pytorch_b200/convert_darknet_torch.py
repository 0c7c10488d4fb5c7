#!/usr/bin/env python
"""convert_darknet_torch -- Darknet `.weights` -> a `.pth` state_dict for this code base (the reference's CLI of the same
name, convert_darknet_torch.py:83-125, without its model-directory / Saver bookkeeping):

    python convert_darknet_torch.py yolo-voc.weights yolo-voc.pth -c config.ini config/darknet/yolo-voc.ini
    python convert_darknet_torch.py model.pth out.weights --reverse -c config.ini config/darknet/yolo-voc.ini

The network is built from the config exactly as detect.py / train.py build it (`model/dnn`, anchors TSV, category list),
so the file is checked against the architecture: leftover or missing floats are reported.  Host code only."""
import argparse
import configparser
import logging
import os

import torch

import model
import utils
from utils import darknet_weights


def build_dnn(config):
    os.chdir(os.path.dirname(os.path.abspath(__file__)))      # config paths (anchors, category) are relative to the package
    category = utils.get_category(config)
    anchors = torch.from_numpy(utils.get_anchors(config)).contiguous()
    dnn = utils.parse_attr(config.get('model', 'dnn'))(model.ConfigChannels(config), anchors, len(category))
    return dnn, anchors


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('src')
    ap.add_argument('dst')
    ap.add_argument('-c', '--config', nargs='+', default=['config.ini'])
    ap.add_argument('-m', '--modify', nargs='+', default=[], help='section/option=value overrides, as in the reference')
    ap.add_argument('--reverse', action='store_true', help='.pth -> .weights')
    args = ap.parse_args(argv)
    src, dst = os.path.abspath(args.src), os.path.abspath(args.dst)
    config = configparser.ConfigParser()
    here = os.path.dirname(os.path.abspath(__file__))
    utils.load_config(config, [p if os.path.isabs(p) else os.path.join(here, p) for p in args.config])
    for cmd in args.modify:
        utils.modify_config(config, cmd)
    dnn, anchors = build_dnn(config)
    if args.reverse:
        state_dict = torch.load(src, map_location='cpu')
        darknet_weights.save_darknet_weights(dst, state_dict, len(anchors))
        return 0
    info = darknet_weights.load_into(dnn, src, len(anchors))
    logging.info('major=%(major)d, minor=%(minor)d, revision=%(revision)d, seen=%(seen)d, %(assigned)d parameters assigned' % info)
    if info['remaining'] > 0:
        logging.warning('%d bytes remaining' % info['remaining'])
    torch.save(dnn.state_dict(), dst)
    return 0


if __name__ == '__main__':
    logging.basicConfig(level=logging.INFO)
    raise SystemExit(main())
