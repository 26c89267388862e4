"""Drop-in mirror of ``code/solver/master.py`` (reference): ``MasterTree`` post-processing of a search --
best policy per scenario and globally, utilities, reward guesses for unexplored (node, scenario) pairs,
pickle save / load.  Host code, as in the reference (O(nodes) work after the search; SURVEY.md section 8f
N1).  Plotting methods are out of scope.
"""
import pickle
from math import log

import numpy as np

from .master_node import MasterNode
from .simulatorTLKT import ACTIONS
from .worker import UCT_COEFF  # bound at import, like the reference (master.py:11; quirk Q4)

ROOT_HASH = hash(tuple([]))


class MasterTree:
    """Final result of a parallel search: ``nodes`` (dictionary of ``MasterNode`` from
    ``master_node.deepcopy_dict``), ``probability`` of each scenario, ``best_policy`` / ``best_nodes_policy``
    per scenario id (``-1`` = global)."""

    def __init__(self, sims, destination, nodes=None, proba=[]):
        num_scenarios = len(sims)
        self.nodes = dict() if nodes is None else nodes
        if len(self.nodes) == 0:
            self.nodes[ROOT_HASH] = MasterNode(num_scenarios, nodehash=ROOT_HASH)
        self.Simulators = sims
        self.numScenarios = num_scenarios
        self.destination = destination
        self.best_policy = dict()
        self.best_nodes_policy = dict()
        if len(proba) != num_scenarios:
            self.probability = np.array([1 / num_scenarios for _ in range(num_scenarios)])
        else:
            self.probability = np.array(proba)

    def get_best_policy(self, verbose=True):
        """Greedy descent from the root for every scenario and for the global tree (master.py:57-84)."""
        for id_scenario in range(-1, len(self.Simulators)):
            if verbose:
                print("Global policy" if id_scenario == -1 else "Policy for scenario {}".format(id_scenario))
            node = self.nodes[ROOT_HASH]
            nodes_policy, policy = [node], []
            while node.children:
                child, action = self.get_best_child(node, idscenario=id_scenario, verbose=verbose)
                if child is None:
                    break
                nodes_policy.append(child)
                policy.append(action)
                node = child
            self.best_policy[id_scenario] = policy
            self.best_nodes_policy[id_scenario] = nodes_policy

    def get_best_child(self, node, idscenario=-1, verbose=False):
        """Child with the largest exploitation value (first maximum wins; values must be > 0)."""
        best_reward, best_action, best_child = 0, None, None
        if idscenario != -1 and all(not child.is_expanded(idscenario) for child in node.children):
            return best_child, best_action
        for child in node.children:
            value, _ = self.get_utility(child, idscenario)
            if value > best_reward:
                best_reward, best_child, best_action = value, child, child.arm
        if verbose and best_child is not None:
            print("Depth {}: best reward = {} for action = {}".format(best_child.depth, best_reward, best_action))
        return best_child, best_action

    def guess_reward(self, node, idscenario):
        """Reward of a node a scenario never explored: the lowest reward for which its father would have been
        expanded next (master.py:117-138)."""
        father = node.parentNode
        if father.is_expanded(idscenario):
            _, exploration = self.get_utility(father, idscenario)
            best_uncle, _ = self.get_best_child(father.parentNode, idscenario)
            value_grd, expl_grd = self.get_utility(best_uncle, idscenario)
            return value_grd + expl_grd - exploration
        return father.guessed_rewards[idscenario]

    def get_utility(self, node, idscenario):
        """(exploitation, exploration) of a node for one scenario, or probability-weighted over the scenarios
        for ``idscenario == -1`` with guessed rewards where a scenario did not expand the node
        (master.py:140-185)."""
        if idscenario != -1 and not node.is_expanded(idscenario):
            return 0, 0
        num_parent = 0
        num_node = 0
        reward_per_action = np.zeros(shape=len(ACTIONS))
        for j in range(len(ACTIONS)):
            if idscenario == -1:
                per_scenario = []
                for i in range(self.numScenarios):
                    if node.is_expanded(i):
                        per_scenario.append(node.rewards[i, j].get_mean())
                        num_node += sum(node.rewards[i, j].h)
                        num_parent += sum(node.parentNode.rewards[i, j].h)
                    else:
                        guess = self.guess_reward(node, i)
                        node.guessed_rewards[i] = guess
                        per_scenario.append(guess)
                reward_per_action[j] = np.dot(per_scenario, self.probability)
            else:
                num_node += sum(node.rewards[idscenario, j].h)
                if node.parentNode is None:
                    num_parent = num_node
                else:
                    num_parent += sum(node.parentNode.rewards[idscenario, j].h)
                reward_per_action[j] = node.rewards[idscenario, j].get_mean()
        value = np.max(reward_per_action)
        exploration = UCT_COEFF * (2 * log(num_parent) / num_node) ** 0.5
        return value, exploration

    def get_depth(self):
        return max(node.depth for node in self.nodes.values())

    def save_tree(self, name, directory="../results/"):
        with open(directory + name + '.pickle', 'wb') as f:
            pickle.dump(self, f)

    @classmethod
    def load_tree(cls, name, directory="../results/"):
        with open(directory + name + '.pickle', 'rb') as f:
            return pickle.load(f)
