"""Stage-by-stage comparison of train.encoder_train_step with the float64 CPU oracle (branch outputs and their gradients)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from conftest import rel_err
    from test_gpu_backward import _oracle_train_step
    import oracle.mmri as om
    from deepinteraction_b200 import mmri, train
    seed = 1570
    cap = {}
    orig = om.DeepInteractionEncoderLayer.forward

    def fwd(self, img_feat, lidar_feat, img_metas, pts_metas, return_parts=False):
        new_img, new_lidar, parts = orig(self, img_feat, lidar_feat, img_metas, pts_metas, return_parts=True)
        B, BN = lidar_feat.shape[0], img_feat.shape[0]
        img5 = img_feat.view(B, -1, *img_feat.shape[1:])
        parts['warped'] = self.P2I_block.Warp(lidar_feat, img5, img_metas, pts_metas).flatten(0, 1).detach()
        for v in parts.values():
            if v.requires_grad:
                v.retain_grad()
        cap.setdefault('layers', []).append(parts)
        return new_img, new_lidar
    om.DeepInteractionEncoderLayer.forward = fwd
    proj = {}
    orig_lcab = om.LocalContextAttentionBlock.forward

    def lcab_fwd(self, target, source, chunk=1):
        hs = []
        rec = proj.setdefault(id(self), {})
        if 'q' in rec:                       # the debug forward of the layer calls Warp only; blocks run once per step
            return orig_lcab(self, target, source, chunk)
        for name, mod in (('q', self.query_project), ('k', self.key_project), ('v', self.value_project)):
            def hook(m, i, o, name=name):
                o.retain_grad()
                rec[name] = o
            hs.append(mod.register_forward_hook(hook))
        out = orig_lcab(self, target, source, chunk)
        out.retain_grad()
        rec['out'] = out
        for h_ in hs:
            h_.remove()
        return out
    om.LocalContextAttentionBlock.forward = lcab_fwd
    train.DEBUG[0] = True
    r64 = _oracle_train_step(seed, torch.float64)
    fr, Gs = r64['fr'], r64['Gs']
    d = torch.device('cuda:0')
    enc = mmri.DeepInteractionEncoder(2, 64, 64, 128)
    enc.load_state_dict(r64['state0'], strict=True)
    enc = enc.to(d).train()
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().to(d)
    pm = {k: (v.to(d) if torch.is_tensor(v) else [p.to(d) for p in v]) for k, v in fr['pts_metas'].items()}
    dbg = {}
    r = train.encoder_train_step(enc, fr['img_feats'].to(d), fr['pts_feats'].to(d), fr['img_metas'], pm,
                             lambda a, b, c: tuple(nhwc(G) for G in Gs), debug=dbg)
    rows = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
    for li, parts in enumerate(cap['layers']):
        for k in ('i2p', 'p2p', 'warped', 'p2i', 'i2i'):
            ours = dbg['fwd'][li][k].reshape(-1, 128).cpu().double()
            ref = rows(parts[k])
            diff = (ours - ref).abs()
            line = 'layer %d %-7s fwd %.2e (rows > 1e-3 of max: %d / %d)' % (li, k, rel_err(ours, ref), int((diff.max(1).values > 1e-3 * ref.abs().max()).sum()), ref.shape[0])
            if k != 'warped' and parts[k].grad is not None:
                g_ref = rows(parts[k].grad)
                g = dbg['bwd'][li][k].reshape(-1, 128).cpu().double()
                line += '   grad %.2e' % rel_err(g, g_ref)
            print(line)
        blk = r64['m'].fusion_blocks[li]
        for name, ob in (('p_iml', blk.P_IML), ('p2i', blk.P2I_block.Local), ('i_iml', blk.I_IML)):
            rec, ours = proj[id(ob)], dbg['fwd'][li]['blocks'][name].dbg
            e = lambda a, b: rel_err(a.cpu().double(), rows(b))
            print('layer %d %-6s q %.1e k %.1e v %.1e | dO %.1e dq %.1e dk %.1e dv %.1e' % (
                li, name, e(ours['q'], rec['q']), e(ours['k'], rec['k']), e(ours['v'], rec['v']), e(ours['dO'], rec['out'].grad),
                e(ours['dq'], rec['q'].grad), e(ours['dk'], rec['k'].grad), e(ours['dv'], rec['v'].grad)))
    errs = sorted(((rel_err(r['grads'][n].cpu().double().view_as(gr), gr), n) for n, gr in r64['grads'].items()
                   if not n.endswith('out_proj.bn.bias')), reverse=True)
    for e_, n in errs[:25]:
        print('%.2e %s' % (e_, n))


if __name__ == '__main__':
    main()
