"""Stand-in for ``nxctmctree.navigation.partition_nodes``.

TEST INFRASTRUCTURE ONLY.  Contract from nxblink/raoteh.py:160-184: yields
pools of nodes that are joined by edges whose branch length or rate is zero
(such nodes must share their state, so their data sets are intersected).
"""
import networkx as nx


def partition_nodes(T, edge_to_blen, edge_to_rate):
    G = nx.Graph()
    G.add_nodes_from(T)
    for edge in T.edges():
        if not edge_to_blen[edge] or not edge_to_rate[edge]:
            G.add_edge(*edge)
    for component in nx.connected_components(G):
        yield list(component)
