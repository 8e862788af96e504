import math


def ranking_evaluation(origin, res, N):
    out = []
    for n in N:
        hits = {u: len(set(origin[u]) & set(i for i, _ in res[u][:n])) for u in origin}
        total = sum(len(origin[u]) for u in origin)
        rec = [hits[u] / len(origin[u]) for u in hits]
        ndcg = 0
        for u in res:
            dcg = sum(1.0 / math.log(p + 2, 2) for p, (i, _) in enumerate(res[u][:n]) if i in origin[u])
            idcg = sum(1.0 / math.log(p + 2, 2) for p in range(min(len(origin[u]), n)))
            ndcg += dcg / idcg
        out += ["Top %d\n" % n, "Hit Ratio:%s\n" % round(sum(hits.values()) / total, 5),
                "Precision:%s\n" % round(sum(hits.values()) / (len(hits) * n), 5), "Recall:%s\n" % round(sum(rec) / len(rec), 5),
                "NDCG:%s\n" % round(ndcg / len(res), 5)]
    return out
