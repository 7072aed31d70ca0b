"""Stage-by-stage encoder comparison (GPU kernels vs CPU oracle) on the medium C=128 scene; prints rel errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepinteraction_b200 import mmri, synth, ops
import oracle.mmri as om


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def main(cloud='dense', views=3, bev=90, in_hw=(224, 400), npts=60000):
    seed = 1700
    torch.manual_seed(seed)
    m = om.DeepInteractionEncoder(2, 32, 48, 128).eval()
    synth.randomize_norm_stats(m, seed)
    fr = synth.make_frame_batch(seed, batch=1, num_views=views, in_hw=in_hw, stride=4, c_img=32, c_pts=48,
                                bev_hw=(bev, bev), n_points=npts, cloud=cloud)
    pil, coors, npts_ = synth.pillarize([p.numpy() for p in fr['pts_metas']['pts']], pillar=108.0 / bev)
    fr['pts_metas'].update(pillars=torch.from_numpy(pil), pillar_coors=torch.from_numpy(coors),
                           pillars_num_points=torch.from_numpy(npts_))
    dev = torch.device('cuda:0')
    enc = mmri.DeepInteractionEncoder(2, 32, 48, 128)
    enc.load_state_dict(m.state_dict(), strict=True)
    enc = enc.to(dev).eval()
    frd = synth.to_device(fr, dev)
    dbg = []
    img_g, p0_g, p1_g = enc.forward_nhwc(frd['img_feats'], frd['pts_feats'], frd['img_metas'], frd['pts_metas'], debug=dbg)
    nchw = lambda t: t.permute(0, 3, 1, 2).cpu()
    with torch.no_grad():
        img = m.shared_conv_img(fr['img_feats'])
        pts = m.shared_conv_pts(fr['pts_feats'])
        print('pillars', len(npts_), 'shared convs done')
        warp_ref, aux = om.BEVWarp()(pts, img.view(1, views, 128, *img.shape[-2:]), fr['img_metas'], fr['pts_metas'], return_aux=True)
        g = enc.last_geometry
        print('sparse equal', torch.equal(g.sparse.cpu(), aux[0]['sparse']), 'max diff', float((g.sparse.cpu() - aux[0]['sparse']).abs().max()))
        print('dense max abs diff', float((g.dense.cpu() - aux[0]['dense']).abs().max()))
        for li, blk in enumerate(m.fusion_blocks):
            new_img, new_pts, parts = blk(img, pts, fr['img_metas'], fr['pts_metas'], return_parts=True)
            d = dbg[li]
            for k in ('i2p', 'p2p', 'p2i', 'i2i'):
                print(f'layer {li} {k:4s} rel err {rel(nchw(d[k]), parts[k]):.2e}')
            wr = om.BEVWarp()(pts, img.view(1, views, 128, *img.shape[-2:]), fr['img_metas'], fr['pts_metas'])[0]
            print(f'layer {li} warp rel err {rel(nchw(d["warped"]), wr):.2e}')
            img, pts = new_img, new_pts
        print('final img', rel(nchw(img_g), img), 'pts', rel(nchw(p1_g), pts))


if __name__ == '__main__':
    main()
    main(cloud='lidar', views=2, bev=36, in_hw=(112, 200), npts=6000)
