"""The drop-in surface driven the way the reference's run_nerf.py drives it (run_nerf.py:352-405, 541-703):
create_nerf(args) -> a few iterations of the training-loop body (render on a ray batch, img2mse, sparsity and TV
terms, RAdam, learning-rate decay) -> checkpoint .tar with the reference's keys -> create_nerf again reloads it ->
render_path on two poses.  Values are checked elsewhere against the oracle; this test pins the call sequence, the
returned structures and the checkpoint contract."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _args(basedir, expname):
    # the Blender/chair configuration of the reference (configs/chair.txt + README.md:20 flags), shrunk
    return argparse.Namespace(
        multires=10, multires_views=4, i_embed=1, i_embed_views=2, use_viewdirs=True, N_importance=32, N_samples=16,
        netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=1024 * 64, lrate=0.01, lrate_decay=10,
        ft_path=None, basedir=str(basedir), expname=expname, no_reload=False, perturb=1., white_bkgd=True,
        raw_noise_std=0., dataset_type='blender', no_ndc=False, lindisp=False, finest_res=128, log2_hashmap_size=14,
        bounding_box=(torch.tensor([-1.5, -1.5, -1.5]), torch.tensor([1.5, 1.5, 1.5])), sparse_loss_weight=1e-10,
        tv_loss_weight=1e-6)


@pytest.fixture(params=[False, True], ids=["default_cpu_tensors", "default_cuda_tensors"])
def default_cuda(request):
    """run_nerf.py:725 switches the global default tensor type to CUDA before train(); the shims must work with
    and without that switch."""
    if request.param:
        torch.set_default_tensor_type('torch.cuda.FloatTensor')
    yield request.param
    if request.param:
        torch.set_default_tensor_type('torch.FloatTensor')


def test_training_loop_body_checkpoint_and_render_path(tmp_path, default_cuda):
    import run_nerf_helpers as H
    from loss import total_variation_loss
    from ray_util import get_rays
    torch.manual_seed(0)
    np.random.seed(0)
    expname = "exp"
    os.makedirs(tmp_path / expname)
    args = _args(tmp_path, expname)
    train_kw, test_kw, start, grad_vars, optimizer = H.create_nerf(args)
    assert start == 0 and {"network_fn", "network_fine", "embed_fn", "network_query_fn"} <= set(train_kw)
    bounds = {"near": 2., "far": 6.}
    train_kw.update(bounds)
    test_kw.update(bounds)
    Himg, Wimg, focal = 16, 16, 20.0
    K = np.array([[focal, 0, 0.5 * Wimg], [0, focal, 0.5 * Himg], [0, 0, 1]])
    pose = torch.tensor([[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 4.0], [0, 0, 0, 1.0]], device=DEV)
    target_img = torch.rand(Himg, Wimg, 3, device=DEV)
    embed_fn = train_kw["embed_fn"]
    N_rand, losses = 64, []
    picks = np.random.choice(Himg * Wimg, size=[N_rand], replace=False)  # one fixed batch: the loss must go down
    for i in range(start + 1, start + 9):                      # run_nerf.py:576-651, no_batching branch
        rays_o, rays_d = get_rays(Himg, Wimg, K, pose[:3, :4])
        coords = torch.stack(torch.meshgrid(torch.linspace(0, Himg - 1, Himg), torch.linspace(0, Wimg - 1, Wimg),
                                            indexing="ij"), -1).reshape(-1, 2)
        sel = coords[picks].long().to(rays_o.device)
        ro, rd = rays_o[sel[:, 0], sel[:, 1]], rays_d[sel[:, 0], sel[:, 1]]
        target_s = target_img[sel[:, 0], sel[:, 1]]
        rgb, depth, acc, extras = H.render(Himg, Wimg, K, chunk=1024, rays=torch.stack([ro, rd], 0), verbose=i < 10,
                                           retraw=True, **train_kw)
        trans = extras['raw'][..., -1]                          # run_nerf.py:614
        assert trans.shape == (N_rand, args.N_samples + args.N_importance)
        optimizer.zero_grad()
        loss = H.img2mse(rgb, target_s)
        psnr = H.mse2psnr(loss)
        loss = loss + H.img2mse(extras["rgb0"], target_s)
        loss = loss + args.sparse_loss_weight * (extras["sparsity_loss"].sum() + extras["sparsity_loss0"].sum())
        n_levels = embed_fn.n_levels
        tv = sum(total_variation_loss(embed_fn.embeddings[l], embed_fn.base_resolution, embed_fn.finest_resolution, l,
                                      embed_fn.log2_hashmap_size, n_levels=n_levels) for l in range(n_levels))
        loss = loss + args.tv_loss_weight * tv
        loss.backward()
        optimizer.step()
        new_lrate = args.lrate * (0.1 ** (i / (args.lrate_decay * 1000)))
        for group in optimizer.param_groups:
            group['lr'] = new_lrate
        losses.append(float(loss.detach()))
        assert rgb.shape == (N_rand, 3) and torch.isfinite(psnr)
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses

    # checkpoint with the reference's keys (run_nerf.py:663-673), reloaded by create_nerf (run_nerf_helpers.py:150-169)
    path = os.path.join(tmp_path, expname, "{:06d}.tar".format(8))
    torch.save({'global_step': 8,
                'network_fn_state_dict': train_kw['network_fn'].state_dict(),
                'network_fine_state_dict': train_kw['network_fine'].state_dict(),
                'embed_fn_state_dict': train_kw['embed_fn'].state_dict(),
                'optimizer_state_dict': optimizer.state_dict()}, path)
    train2, test2, start2, _, opt2 = H.create_nerf(args)
    assert start2 == 8
    for a, b in zip(train_kw['embed_fn'].parameters(), train2['embed_fn'].parameters()):
        assert torch.equal(a, b)
    for a, b in zip(train_kw['network_fine'].parameters(), train2['network_fine'].parameters()):
        assert torch.equal(a, b)
    s1, s2 = optimizer.state_dict()['state'], opt2.state_dict()['state']
    assert s1.keys() == s2.keys() and all(torch.equal(s1[k]['exp_avg'], s2[k]['exp_avg']) for k in s1)
    assert list(train2['embed_fn'].state_dict().keys())[0] == "embeddings.0.weight"

    # test-time rendering of a short camera path (run_nerf.py:397, run_nerf_helpers.py:386-459)
    test2.update(bounds)
    poses = torch.stack([pose, pose.clone()])
    poses[1, 0, 3] = 0.3
    with torch.no_grad():
        rgbs, depths = H.render_path(poses, (Himg, Wimg, focal), K, 1024, test2, gt_imgs=None, savedir=None)
        again, _ = H.render_path(poses[:1], (Himg, Wimg, focal), K, 1024, test2)
    assert rgbs.shape == (2, Himg, Wimg, 3) and depths.shape == (2, Himg, Wimg)
    assert np.isfinite(rgbs).all() and rgbs.min() >= -1e-5 and rgbs.max() <= 1 + 1e-5
    np.testing.assert_array_equal(rgbs[0], again[0])           # perturb = 0 at test time: deterministic
    assert H.to8b(rgbs[0]).dtype == np.uint8
