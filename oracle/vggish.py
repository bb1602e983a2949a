"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference's VGGish feature stack.

Follows ``VGGish.forward_feat`` = ``VGG.forward_feat`` (models/vggish.py:87-90) over ``make_layers()``
(models/vggish.py:93-103): conv3x3(1->64)+ReLU, pool, conv(64->128)+ReLU, pool, conv(128->256)+ReLU,
conv(256->256)+ReLU, pool, conv(256->512)+ReLU, conv(512->512)+ReLU, pool (all pools 2x2 stride 2), as it is called by
``VideoSaliencyModel.forward_vggish`` (models/diff_model.py:70-81) on ``audio.view(-1, 1, 112, 192)``.
Only tests/ may import this module.  Pinned against the reference class in tests/test_oracle_vs_reference.py.
"""
import torch.nn.functional as F

CONV_IDX = (0, 3, 6, 8, 11, 13)            # positions of the Conv2d modules inside ``features``
POOL_AFTER = (0, 3, 8, 13)                 # convs followed by MaxPool2d(2, 2)


def forward_feat(sd, x):
    """x [(B*T), 1, H, W] fp32 -> [(B*T), 512, H/16, W/16]."""
    for i in CONV_IDX:
        x = F.relu(F.conv2d(x, sd["features.%d.weight" % i], sd["features.%d.bias" % i], padding=1))
        if i in POOL_AFTER:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
    return x
