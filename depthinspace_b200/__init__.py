"""depthinspace_b200 -- B200 (sm_100a) implementation of DepthInSpace's self-supervision hot path.

LCN -> disparity / optical-flow warp -> k x k block-window photometric loss -> edge-aware
smoothness, forward and backward, as hand-written CUDA behind a C-ABI (include/dis_b200.h,
lib/libdis_b200.so) and behind the reference's own Python surface:

    depthinspace_b200.ext_functions         <-> reference model/ext_functions.py
    depthinspace_b200.networks              <-> reference model/networks.py (LCN, losses, Sobel)
    depthinspace_b200.multi_frame_networks  <-> reference model/multi_frame_networks.py (warp)
    depthinspace_b200.losses                <-> loss assembly of model/{single,multi}_frame_worker.py
    depthinspace_b200.torchext/             <-> drop-in `ext_cuda` / `ext_cpu` modules (CTD_DIR)
"""
__version__ = "0.1.0"
