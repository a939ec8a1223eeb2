"""Agent registry (reference: src/modules/agents/__init__.py:1-13).  The four entity-attention agents share one
implementation; `imagine_*` only changes which mask copies the controller requests."""
from ..nets import EntityAttnAgent

REGISTRY = {
    "entity_attend_rnn": EntityAttnAgent,
    "imagine_entity_attend_rnn": EntityAttnAgent,
    "entity_attend_ff": EntityAttnAgent,
    "imagine_entity_attend_ff": EntityAttnAgent,
}
