"""Your training script.  Minimal data-parallel loop on the b200-ddl runtime:

    from distributeddeeplearning_b200 import models, ops
    from distributeddeeplearning_b200.parallel import DistributedOptimizer, dist

    dist.init()
    model = models.get_model("resnet50").cuda()
    opt = DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.0125 * dist.size(), momentum=0.9),
                               named_parameters=model.named_parameters())
    for data, target in loader:
        opt.zero_grad(); loss = ops.softmax_cross_entropy(model(data), target); loss.backward(); opt.step()
"""
