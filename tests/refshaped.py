"""Test-local STOCK torch.nn models with the shape of the reference's model files -- operators bound through
`import torch.nn as nn` / `import torch.nn.functional as F` at import, forward()s that call F.upsample, F.dropout3d,
torch.cat, in-place ReLU and a `.view` Flatten -- so that `nn.convert()` is exercised on what a reference user hands it
(unet3d.py:20-79, cnn_model.py:8-10,43-101).  Written for the tests; not a copy of the reference files."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Down(nn.Module):
    def __init__(self, cin, cout, first=False):
        super().__init__()
        self.first = first
        self.pool = nn.MaxPool3d(2, 2)
        self.c1, self.n1 = nn.Conv3d(cin, cout, 3, 1, 1, bias=False), nn.BatchNorm3d(cout)
        self.c2, self.n2 = nn.Conv3d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm3d(cout)
        self.c3, self.n3 = nn.Conv3d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm3d(cout)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        if not self.first:
            x = self.pool(x)
        x = self.n1(self.c1(x))
        y = self.relu(self.n2(self.c2(x)))
        y = F.dropout3d(y, 0.5)                 # result overwritten below, as in the reference's encoder stage
        y = self.n3(self.c3(x))
        return self.relu(x + y)


class Up(nn.Module):
    def __init__(self, planes, first=False):
        super().__init__()
        self.first = first
        if not first:
            self.c1, self.n1 = nn.Conv3d(2 * planes, planes, 3, 1, 1, bias=False), nn.BatchNorm3d(planes)
        self.c2, self.n2 = nn.Conv3d(planes, planes // 2, 1, 1, 0, bias=False), nn.BatchNorm3d(planes // 2)
        self.c3, self.n3 = nn.Conv3d(planes, planes, 3, 1, 1, bias=False), nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x, skip):
        if not self.first:
            x = self.relu(self.n1(self.c1(x)))
        y = F.upsample(x, scale_factor=2, mode="trilinear", align_corners=False)
        y = self.relu(self.n2(self.c2(y)))
        y = torch.cat([skip, y], 1)
        return self.relu(self.n3(self.c3(y)))


class RefShapedUNet(nn.Module):
    def __init__(self, n=16, classes=2):
        super().__init__()
        self.d1, self.d2, self.d3 = Down(1, n, True), Down(n, 2 * n), Down(2 * n, 4 * n)
        self.u2, self.u1 = Up(4 * n, first=True), Up(2 * n)
        self.bridge = nn.Conv3d(4 * n, 4 * n, 1, bias=False)
        self.head2, self.head1 = nn.Conv3d(4 * n, classes, 1), nn.Conv3d(2 * n, classes, 1)
        self.upsample = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False)

    def forward(self, x):
        x1 = self.d1(x)
        x2 = self.d2(x1)
        x3 = self.bridge(self.d3(x2))
        y2 = self.u2(x3, x2)
        y1 = self.u1(y2, x1)
        return self.head1(y1) + self.upsample(self.head2(y2))


class Flatten(nn.Module):
    def forward(self, input):
        return input.view(input.size(0), -1)


class RefShapedClassifier(nn.Module):
    def __init__(self, n=8, shape=(16, 16, 16)):
        super().__init__()
        self.model = nn.Sequential()
        self.model.add_module("conv3d_1", nn.Conv3d(1, n, kernel_size=3, padding=1, stride=2))
        self.model.add_module("batch_norm_1", nn.BatchNorm3d(n))
        self.model.add_module("activation_1", nn.ReLU(inplace=True))
        self.model.add_module("conv3d_2", nn.Conv3d(n, 2 * n, kernel_size=3, padding=1))
        self.model.add_module("batch_norm_2", nn.BatchNorm3d(2 * n))
        self.model.add_module("activation_2", nn.LeakyReLU())
        self.model.add_module("max_pool3d_1", nn.MaxPool3d(kernel_size=2))
        self.model.add_module("flatten_1", Flatten())
        self.model.add_module("fully_conn_1", nn.Linear(2 * n * (shape[0] // 4) * (shape[1] // 4) * (shape[2] // 4), 16))
        self.model.add_module("batch_norm_9", nn.BatchNorm1d(16))
        self.model.add_module("fully_conn_2", nn.Linear(16, 2))

    def forward(self, x):
        return self.model(x)


# ------------------------------------------------------------------------------------------------ full-size stock restatement of unet3d.Unet
def _norm(planes, kind):
    if kind == "bn":
        return nn.BatchNorm3d(planes)
    if kind == "gn":
        return nn.GroupNorm(4, planes)
    if kind == "in":
        return nn.InstanceNorm3d(planes)
    raise ValueError(kind)


class StockConvD(nn.Module):
    def __init__(self, cin, cout, dropout, norm, first=False):
        super().__init__()
        self.first, self.dropout = first, dropout
        self.maxpool = nn.MaxPool3d(2, 2)
        self.relu = nn.ReLU(inplace=True)
        self.conv1, self.bn1 = nn.Conv3d(cin, cout, 3, 1, 1, bias=False), _norm(cout, norm)
        self.conv2, self.bn2 = nn.Conv3d(cout, cout, 3, 1, 1, bias=False), _norm(cout, norm)
        self.conv3, self.bn3 = nn.Conv3d(cout, cout, 3, 1, 1, bias=False), _norm(cout, norm)

    def forward(self, x):
        if not self.first:
            x = self.maxpool(x)
        x = self.bn1(self.conv1(x))
        y = self.relu(self.bn2(self.conv2(x)))
        if self.dropout > 0:
            y = F.dropout3d(y, self.dropout)
        y = self.bn3(self.conv3(x))
        return self.relu(x + y)


class StockConvU(nn.Module):
    def __init__(self, planes, norm, first=False):
        super().__init__()
        self.first = first
        if not first:
            self.conv1, self.bn1 = nn.Conv3d(2 * planes, planes, 3, 1, 1, bias=False), _norm(planes, norm)
        self.conv2, self.bn2 = nn.Conv3d(planes, planes // 2, 1, 1, 0, bias=False), _norm(planes // 2, norm)
        self.conv3, self.bn3 = nn.Conv3d(planes, planes, 3, 1, 1, bias=False), _norm(planes, norm)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x, prev):
        if not self.first:
            x = self.relu(self.bn1(self.conv1(x)))
        y = F.upsample(x, scale_factor=2, mode="trilinear", align_corners=False)
        y = self.relu(self.bn2(self.conv2(y)))
        y = torch.cat([prev, y], 1)
        return self.relu(self.bn3(self.conv3(y)))


class StockUnet3d(nn.Module):
    """A stock-torch.nn model with the operator graph, attribute names and state_dict keys of the reference's
    segmentation/models/unet3d.py `Unet` (written for the tests and the bench's drop-in leg, where /root/reference does not exist):
    what a reference user holds before calling `nn.convert()`."""

    def __init__(self, c=4, n=16, dropout=0.5, norm="gn", num_classes=5):
        super().__init__()
        self.upsample = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False)
        w = [c, n, 2 * n, 4 * n, 8 * n, 16 * n]
        for i in range(1, 6):
            setattr(self, f"convd{i}", StockConvD(w[i - 1], w[i], dropout, norm, first=(i == 1)))
        self.convu4 = StockConvU(16 * n, norm, True)
        self.convu3, self.convu2, self.convu1 = StockConvU(8 * n, norm), StockConvU(4 * n, norm), StockConvU(2 * n, norm)
        self.seg3, self.seg2, self.seg1 = nn.Conv3d(8 * n, num_classes, 1), nn.Conv3d(4 * n, num_classes, 1), nn.Conv3d(2 * n, num_classes, 1)

    def forward(self, x):
        x1 = self.convd1(x)
        x2 = self.convd2(x1)
        x3 = self.convd3(x2)
        x4 = self.convd4(x3)
        x5 = self.convd5(x4)
        y4 = self.convu4(x5, x4)
        y3 = self.convu3(y4, x3)
        y2 = self.convu2(y3, x2)
        y1 = self.convu1(y2, x1)
        y3 = self.seg3(y3)
        y2 = self.seg2(y2) + self.upsample(y3)
        return self.seg1(y1) + self.upsample(y2)
