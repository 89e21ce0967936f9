"""construct_feed_dict_train / construct_feed_dict with the reference's
signatures (preprocessing.py:32-50): map the numpy batches to placeholders."""


def construct_feed_dict_train(features, spatial, adj, rel, adj_truth, feature_truth_batch, spatial_truth_batch,
                              rel_truth_batch, placeholders):
    feed_dict = dict()
    feed_dict.update({placeholders['features']: features})
    feed_dict.update({placeholders['adj']: adj})
    feed_dict.update({placeholders['spatial']: spatial})
    feed_dict.update({placeholders['rel']: rel})
    feed_dict.update({placeholders['adj_truth']: adj_truth})
    feed_dict.update({placeholders['feature_truth']: feature_truth_batch})
    feed_dict.update({placeholders['spatial_truth']: spatial_truth_batch})
    feed_dict.update({placeholders['rel_truth']: rel_truth_batch})
    return feed_dict


def construct_feed_dict(features, spatial, adj, rel, placeholders):
    feed_dict = dict()
    feed_dict.update({placeholders['features']: features})
    feed_dict.update({placeholders['adj']: adj})
    feed_dict.update({placeholders['spatial']: spatial})
    feed_dict.update({placeholders['rel']: rel})
    return feed_dict
