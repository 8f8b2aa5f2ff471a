// Prompt + generated tokens; clicking one selects it in every view.  Keeps the last runs for A/B comparison of
// perturbation experiments (what the reference shows as the "Result" text boxes).
import {Component, h, panel} from "./base.js";
import {state, set} from "../store.js";
export class TokenStrip extends Component {
  constructor() { super(); this.box = h("div", {class:"tokens"}); this.text = h("pre", {style:"white-space:pre-wrap;margin:6px 0 0;color:var(--mut)"}); this.hist = h("div", {});
    this.el = panel("Tokens", this.box, this.text, this.hist); this.watch("tokens", "history"); }
  render() {
    this.box.innerHTML = ""; const np = state.promptTokens.length;
    state.promptTokens.concat(state.genTokens).forEach((t, i) => { const s = h("span", {class:(i >= np ? "gen " : "") + (i === state.selToken ? "sel" : "")}, String(t));
      s.onclick = () => set({selToken: i === state.selToken ? -1 : i}, "tokens"); this.box.appendChild(s); });
    this.text.textContent = (state.text || []).join("\n");
    this.hist.innerHTML = ""; state.history.slice(1, 4).forEach(r => this.hist.appendChild(h("div", {style:"color:var(--mut);font-size:11px"}, `${r.when}: ${(r.text && r.text[0]) || r.tokens.join(" ")}`.slice(0, 200))));
  }
}
